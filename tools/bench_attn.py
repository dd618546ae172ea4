"""Micro-benchmark of fresco_attn_fwd at the BASELINE shapes (CUDA events, inputs >> L2 are not needed here: K/V of one
(chunk, head) are re-read by 32 query tiles by design; Q/K/V/O together are 100-250 MB per launch).

    python tools/bench_attn.py [--sdpa] VARIANT [VARIANT ...]
    VARIANT = NAME=VALUE[,NAME=VALUE...]   e.g.  FRESCO_ATTN_WIDE=1,FRESCO_ATTN_POLY=4     ("default" = built-in defaults)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fresco_b200 import _lib, ops  # noqa: E402

SHAPES = [  # B, L, Lk, heads, d, q_per_kv
    (16, 4096, 15587, 8, 40, 8),      # level B, cross-frame (bench.py's synthetic occlusion)
    (16, 4096, 4096, 8, 40, 1),       # level B, spatial-guided
    (16, 1024, 3897, 8, 80, 8),       # level A, cross-frame
    (16, 1024, 1024, 8, 80, 1),       # level A, spatial-guided
    (32, 1024, 1024, 1, 128, 1),      # GMFlow-like single head, d = 128
    (16, 2048, 2048, 5, 64, 1),       # head_dim 64 (not a FRESCO shape: picks the default kernel for it)
]
OPTS = ("FRESCO_ATTN_WIDE", "FRESCO_ATTN_POLY", "FRESCO_ATTN_ROWSUM")


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    do_sdpa = "--sdpa" in sys.argv
    data = {}
    for (B, L, Lk, H, d, qpk) in SHAPES:
        C = H * d
        q = torch.randn(B, L, C, device="cuda").half()
        k = torch.randn(B // qpk, Lk, C, device="cuda").half()
        v = torch.randn(B // qpk, Lk, C, device="cuda").half()
        data[(B, L, Lk, H, d, qpk)] = (q, k, v, torch.empty_like(q))
    for variant in args or ["default"]:
        kv = {} if variant == "default" else dict(a.split("=") for a in variant.split(",") if a)
        for name in OPTS:
            _lib.set_option(name, int(kv.get(name, -1)))
        row = {"variant": variant}
        for shp, (q, k, v, out) in data.items():
            B, L, Lk, H, d, qpk = shp
            ms = timeit(lambda: ops.attn_fwd(q, k, v, H, qpk, out=out))
            row["d%d_L%d_Lk%d" % (d, L, Lk)] = [round(ms, 4), round(4.0 * B * L * Lk * H * d / ms / 1e9, 1)]
        print(json.dumps(row), flush=True)
    if do_sdpa:
        row = {"variant": "torch SDPA (K/V broadcast as the reference does)"}
        for shp, (q, k, v, out) in data.items():
            B, L, Lk, H, d, qpk = shp
            kk = k.repeat_interleave(qpk, 0).view(B, Lk, H, d).transpose(1, 2)
            vv = v.repeat_interleave(qpk, 0).view(B, Lk, H, d).transpose(1, 2)
            qq = q.view(B, L, H, d).transpose(1, 2)
            ms = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qq, kk, vv))
            row["d%d_L%d_Lk%d" % (d, L, Lk)] = [round(ms, 4), round(4.0 * B * L * Lk * H * d / ms / 1e9, 1)]
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
