"""Time fresco_attn_fwd on the config2 head_dim-40 shapes under different values of one environment knob:
    python tools/env_sweep.py FRESCO_ATTN_POLY=0 FRESCO_ATTN_POLY=4,FRESCO_ATTN_ROWSUM=1"""
import json, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
code = r'''
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath("%s"))))
from fresco_b200 import ops
res = []
for (B,L,Lk,H,d,qpk) in [(16,4096,11874,8,40,8), (16,4096,4096,8,40,1)]:
    q=torch.randn(B,L,H*d,device="cuda").half(); k=torch.randn(B//qpk,Lk,H*d,device="cuda").half(); v=torch.randn_like(k); out=torch.empty_like(q)
    for _ in range(3): ops.attn_fwd(q,k,v,H,qpk,out=out)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.attn_fwd(q,k,v,H,qpk,out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/10
    res.append((round(ms,4), round(4.0*B*L*Lk*H*d/ms/1e9,1)))
print(json.dumps(res))
''' % os.path.join(HERE, "x")
for s in sys.argv[1:]:                          # each argument: NAME=VALUE[,NAME=VALUE...]
    kv = dict(a.split("=") for a in s.split(",") if a)
    env = dict(os.environ, **kv)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(json.dumps({"env": s, "ms_tflops": r.stdout.strip() or r.stderr[-300:]}), flush=True)
