"""Numerics of the twin attention kernel, restated in torch on the CPU (fp32 arithmetic where the kernel uses fp32, fp16
rounding where it stores fp16): lazy running max with the 2^8 threshold, p = exp2(s * scale * log2e - m) by MUFU or -- for
every POLY-th pair -- by the Cody-Waite + degree-3 polynomial of exp2_poly_x2 (attn_tcgen05.cu), P rounded to fp16, O
and the row sum accumulated in fp32 from the SAME rounded P (the row sum comes from the tensor core, through a ones
column), one division at the end.  The GPU tests check the kernel against fp32 at 2e-3 * max|y|; this checks that the
ALGORITHM meets that bound with margin on inputs picked to stress it (peaky rows that trigger the rescale, rows whose
scores sink far below the running max, ragged tails, the diagonal bias), so a failure on the GPU points at the
implementation, not at the design."""
import math

import pytest
import torch

LOG2E = 1.4426950408889634


def exp2_poly(t):
    """exp2_poly_x2 of attn_tcgen05.cu in fp32: clamp, round-to-nearest by the 1.5 * 2^23 trick, degree-3 minimax on
    [-0.5, 0.5], exponent patched in with an integer add."""
    t = torch.clamp(t, min=-126.0)
    magic = torch.tensor(12582912.0)
    r = t + magic
    n = r - magic
    f = t - n
    q = torch.tensor(0.05517164617776871) * f + torch.tensor(0.2426111251115799)
    q = q * f + torch.tensor(0.6932609677314758)
    q = q * f + torch.tensor(0.9999280571937561)
    bits = q.view(torch.int32) + (r.view(torch.int32) << 23)
    return bits.view(torch.float32)


def twin_model(q, k, v, scale, poly, split, diag_bias=0.0):
    """q [Lq, d], k, v [Lk, d] fp16 (one head).  Returns fp16 [Lq, d]."""
    Lq, d = q.shape
    Lk = k.shape[0]
    qf, kf, vf = q.float(), k.float(), v.float()
    scale_log2, bias_log2 = scale * LOG2E, diag_bias * LOG2E
    O = torch.zeros(Lq, d)
    l = torch.zeros(Lq)
    m_run = torch.full((Lq,), -math.inf)
    cols_per_thread = 128 // split
    for j in range((Lk + 127) // 128):
        lo, hi = j * 128, min(Lk, j * 128 + 128)
        s = torch.full((Lq, 128), -math.inf)
        s[:, :hi - lo] = qf @ kf[lo:hi].t()
        if diag_bias:
            rows = torch.arange(Lq)
            inside = (rows >= lo) & (rows < hi)
            s[rows[inside], rows[inside] - lo] += bias_log2 / scale_log2
        m_tile = s.max(dim=1).values * scale_log2
        if j == 0:
            m_run = m_tile.clone()
        else:
            need = m_tile > m_run + 8.0
            alpha = torch.where(need, torch.exp2(m_run - m_tile), torch.ones(Lq))
            O *= alpha[:, None]
            l *= alpha
            m_run = torch.where(need, m_tile, m_run)
        t = s * scale_log2 - m_run[:, None]
        p = torch.exp2(t)
        if poly:
            c = torch.arange(128) % cols_per_thread                  # column inside a thread's share of the tile
            use_poly = ((c >> 1) % poly) == (poly - 1)
            p[:, use_poly] = exp2_poly(t[:, use_poly])
        p16 = p.half().float()
        O += p16[:, :hi - lo] @ vf[lo:hi]
        l += p16.sum(dim=1)
    return (O / l[:, None]).half()


def reference(q, k, v, scale, diag_bias=0.0):
    s = q.double() @ k.double().t() * scale
    if diag_bias:
        n = min(s.shape)
        s[torch.arange(n), torch.arange(n)] += diag_bias
    return torch.softmax(s, -1) @ v.double()


@pytest.mark.parametrize("poly,split", [(0, 1), (4, 1), (8, 2), (2, 2)])
@pytest.mark.parametrize("d", [40, 80])
def test_twin_algorithm_meets_the_stated_tolerance(poly, split, d):
    g = torch.Generator().manual_seed(100 * d + 10 * poly + split)
    worst = 0.0
    for (Lq, Lk, gain, scale, bias) in [(96, 1000, 1.0, None, 0.0), (64, 333, 4.0, None, 0.0), (64, 130, 8.0, None, 0.0),
                                        (200, 200, 4.0, 0.2, 1.5), (32, 4096 + 77, 1.0, None, 0.0)]:
        q = (torch.randn(Lq, d, generator=g) * gain).half()
        k = torch.randn(Lk, d, generator=g).half()
        v = torch.randn(Lk, d, generator=g).half()
        # a few rows whose late keys dominate (forces the lazy rescale) and rows of nearly equal scores
        q[0] = (k[Lk - 3].float() * 6).half()
        q[1] = 0
        sc = (scale if scale is not None else 1.0) / math.sqrt(d)
        got = twin_model(q, k, v, sc, poly, split, bias).double()
        want = reference(q, k, v, sc, bias)
        err = (got - want).abs().max().item() / max(1.0, want.abs().max().item())
        worst = max(worst, err)
    assert worst < 1e-3, worst                                     # the GPU tests allow 2e-3


def test_polynomial_exp2_error_and_range():
    t = torch.linspace(-100.0, 9.0, 200001)                         # (P is fp16: anything below 2^-25 rounds to zero anyway)
    ref = torch.exp2(t.double())
    rel = ((exp2_poly(t).double() - ref) / ref).abs().max().item()
    assert rel < 1e-4, rel                                          # 7.5e-5 by design: inside the fp16 rounding of P
    edge = exp2_poly(torch.tensor([-1e30, -140.0, -126.0, 0.0, 8.0]))
    assert torch.isfinite(edge).all() and (edge[:3] < 2e-38).all() and abs(edge[3].item() - 1.0) < 1e-4
