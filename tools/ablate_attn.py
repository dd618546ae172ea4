"""Ablation of the pipelined attention kernel (FRESCO_ATTN_WIDE=0, built with FRESCO_NVCC_EXTRA=-DFRESCO_ATTN_ABLATE_BUILD;
results are wrong on purpose; timing only): which stage sets the pace?"""
import json, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
code = r'''
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath("%s"))))
from fresco_b200 import ops
B,L,Lk,H,d,qpk = 16,4096,11874,8,40,8
q=torch.randn(B,L,H*d,device="cuda").half(); k=torch.randn(B//qpk,Lk,H*d,device="cuda").half(); v=torch.randn_like(k); out=torch.empty_like(q)
for _ in range(3): ops.attn_fwd(q,k,v,H,qpk,out=out)
torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.attn_fwd(q,k,v,H,qpk,out=out)
e1.record(); torch.cuda.synchronize()
print(round(e0.elapsed_time(e1)/10,4))
''' % os.path.join(HERE, "x")
names = {0: "baseline", 1: "no exp2 (MUFU)", 2: "no S load (TMEM->RF)", 4: "no P store (RF->TMEM)", 8: "no PV MMA",
         32: "no QK MMA", 64: "no K/V TMA after the first ring fill", 40: "no MMAs at all", 104: "no MMAs, no TMA",
         23: "no softmax work (exp, S load, P store, max)", 127: "barrier skeleton only"}
for a, n in names.items():
    env = dict(os.environ, FRESCO_ATTN_ABLATE=str(a), FRESCO_ATTN_WIDE="0")      # the pipelined kernel carries the ablation switches
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(json.dumps({"ablate": a, "what": n, "ms": r.stdout.strip() or r.stderr[-300:]}))
