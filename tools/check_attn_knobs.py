"""Parity of every attention tuning-knob combination against an fp32 torch reference (GPU): ragged shapes with a peaky
softmax (lazy-rescale path) and the scale + diagonal-bias path.  Prints error / tolerance per case."""
import os, sys, itertools, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fresco_b200 import ops


def ref_sdpa(q, k, v, H, qpk, scale=None, bias=0.0):
    B, L, C = q.shape
    d = C // H
    scale = scale or 1 / math.sqrt(d)
    kk = k.repeat_interleave(qpk, 0).view(B, -1, H, d).transpose(1, 2).float()
    vv = v.repeat_interleave(qpk, 0).view(B, -1, H, d).transpose(1, 2).float()
    s = q.view(B, L, H, d).transpose(1, 2).float() @ kk.transpose(-1, -2) * scale
    if bias:
        s = s + torch.eye(L, kk.shape[2], device=q.device) * bias
    return (torch.softmax(s, -1) @ vv).transpose(1, 2).reshape(B, L, C)


WIDE = [0, 1] if "--wide" in sys.argv else [0]      # the wide kernel is experimental: only on request
for d in (40, 80):
    torch.manual_seed(17 + d)
    B, L, Lk, H, qpk = 4, 300, 1000, 2, 2
    for gain in (1.0, 4.0):
        q = (torch.randn(B, L, H * d, device="cuda") * gain).half()
        k = torch.randn(B // qpk, Lk, H * d, device="cuda").half()
        v = torch.randn_like(k)
        ref = ref_sdpa(q, k, v, H, qpk)
        qs, ks, vs = q[:2, :256].contiguous(), k[:1, :256].repeat(2, 1, 1), v[:1, :256].repeat(2, 1, 1)
        ref2 = ref_sdpa(qs, ks, vs, H, 1, 0.2 / math.sqrt(d), 1.5)
        for wide, narrow, poly, rowsum in itertools.product(WIDE, [0, 4], [0, 4], [0, 1]):
            if ((narrow or wide) and (poly or rowsum)) or (narrow and wide) or (d == 80 and (narrow or poly or rowsum or wide)):
                continue
            os.environ.update(FRESCO_ATTN_WIDE=str(wide), FRESCO_ATTN_NARROW=str(narrow), FRESCO_ATTN_POLY=str(poly),
                              FRESCO_ATTN_ROWSUM=str(rowsum))
            e1 = (ops.attn_fwd(q, k, v, H, qpk).float() - ref).abs()
            e2 = (ops.attn_fwd(qs, ks, vs, H, 1, softmax_scale=0.2 / math.sqrt(d), diag_bias=1.5).float() - ref2).abs().max().item()
            bad = (e1 > 2e-3 * max(1.0, ref.abs().max().item())).nonzero()
            print(f"d={d} gain={gain} wide={wide} narrow={narrow} poly={poly} rowsum={rowsum}: err {e1.max().item():.2e} / tol "
                  f"{2e-3 * max(1.0, ref.abs().max().item()):.2e} (bad {len(bad)}, first {bad[:1].tolist()}); "
                  f"bias case err {e2:.2e} / tol {2e-3 * ref2.abs().max().item():.2e}")
