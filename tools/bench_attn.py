"""Micro-benchmark of fresco_attn_fwd at the BASELINE shapes (CUDA events, L2-sized inputs)."""
import json
import math
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fresco_b200 import ops  # noqa: E402


def run(B, L, Lk, heads, d, qpk, iters=10):
    C = heads * d
    q = torch.randn(B, L, C, device="cuda").half()
    k = torch.randn(B // qpk, Lk, C, device="cuda").half()
    v = torch.randn(B // qpk, Lk, C, device="cuda").half()
    out = torch.empty_like(q)
    for _ in range(3):
        ops.attn_fwd(q, k, v, heads, qpk, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.attn_fwd(q, k, v, heads, qpk, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 4.0 * B * L * Lk * C
    # torch SDPA on the same shapes (the reference's GPU path), K/V broadcast as the reference does
    kk = k.repeat_interleave(qpk, 0).view(B, Lk, heads, d).transpose(1, 2)
    vv = v.repeat_interleave(qpk, 0).view(B, Lk, heads, d).transpose(1, 2)
    qq = q.view(B, L, heads, d).transpose(1, 2)
    for _ in range(2):
        torch.nn.functional.scaled_dot_product_attention(qq, kk, vv)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.nn.functional.scaled_dot_product_attention(qq, kk, vv)
    e1.record()
    torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / iters
    return {"B": B, "L": L, "Lk": Lk, "heads": heads, "d": d, "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1),
            "torch_sdpa_ms": round(ms_t, 4), "torch_sdpa_tflops": round(flops / ms_t / 1e9, 1)}


if __name__ == "__main__":
    for cfg in [(16, 4096, 11874, 8, 40, 8), (16, 4096, 4096, 8, 40, 1), (16, 1024, 2963, 8, 80, 8),
                (16, 1024, 1024, 8, 80, 1), (16, 4096, 15587, 8, 40, 8)]:
        print(json.dumps(run(*cfg)))
