"""Per-stage SM-clock stamps of one CTA of the PIPELINED attention kernel (needs a -DFRESCO_ATTN_TRACE build:
FRESCO_NVCC_EXTRA=-DFRESCO_ATTN_TRACE; the stamps are compiled into fresco_attn_kernel only)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fresco_b200 import ops, _lib

_lib.set_option("FRESCO_ATTN_WIDE", 0)

d = int(sys.argv[1]) if len(sys.argv) > 1 else 40
B, L, Lk, H, qpk = (16, 4096, 11874, 8, 8) if d == 40 else (16, 1024, 11874, 8, 8)
q = torch.randn(B, L, H * d, device="cuda").half()
k = torch.randn(B // qpk, Lk, H * d, device="cuda").half()
v = torch.randn_like(k)
out = torch.empty_like(q)
for _ in range(3):
    ops.attn_fwd(q, k, v, H, qpk, out=out)
torch.cuda.synchronize()
lib = _lib.lib()
buf = (ctypes.c_longlong * (32 * 16))()
assert lib.fresco_debug_attn_trace(buf) == 0
rows = [[buf[t * 16 + s] for s in range(13)] for t in range(32)]
names = ["top", "S_rdy", "ld_done", "probes", "max", "exp+pack", "st_iss", "st_wait", "arriveP", "QK:c_seen", "QK:issued",
         "PV:p_seen", "PV:issued"]
base = rows[0][0]
print("tile " + " ".join(f"{n:>9}" for n in names) + "   (softmax stamps: delta to previous stamp; issuer stamps: offset from this tile's top)")
for t, r in enumerate(rows):
    cells = [f"{r[0] - base:9d}"]
    for s in range(1, 9):
        cells.append(f"{r[s] - r[s - 1]:9d}")
    for s in range(9, 13):
        cells.append(f"{r[s] - r[0]:9d}")
    print(f"{t + 64:4d} " + " ".join(cells))
print("mean clk per tile:", (rows[31][0] - rows[0][0]) / 31)
