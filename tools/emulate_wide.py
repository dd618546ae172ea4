"""numpy model of the arithmetic of fresco_attn_wide_kernel (two independent key halves per row with their own lazy
running max / row sum / accumulator, fp16 P, merge in the epilogue), checked against a plain softmax.  It mirrors the
kernel's control flow statement by statement (tile loop, lazy threshold 8, masked-half handling, epilogue merge), so a
logic slip in the kernel's scheme shows up here without a GPU.  It says nothing about the PTX."""
import numpy as np


def wide_row_block(S, V, kv_len, scale_log2):
    """S [rows, Lk_pad] raw scores (fp32), V [Lk_pad, d]; returns [rows, d]."""
    rows, Lp = S.shape
    d = V.shape[1]
    n_tiles = Lp // 64
    m_run = np.full((2, rows), -np.inf, np.float32)
    l = np.zeros((2, rows), np.float32)
    O = np.zeros((2, rows, d), np.float32)
    for i in range(n_tiles):
        for h in range(2):
            c0 = i * 64 + 32 * h
            r = S[:, c0:c0 + 32].copy()
            cols = c0 + np.arange(32)
            r[:, cols >= kv_len] = -np.inf
            m_tile = r.max(1) * scale_log2
            if i == 0:
                m_run[h] = m_tile
            else:
                need = m_tile > m_run[h] + 8.0
                with np.errstate(invalid="ignore"):
                    alpha = np.where(need, np.exp2(m_run[h] - m_tile), 1.0).astype(np.float32)
                l[h] *= alpha
                O[h] *= alpha[:, None]
                m_run[h] = np.where(need, m_tile, m_run[h])
            neg_m = np.where(np.isneginf(m_run[h]), 0.0, -m_run[h]).astype(np.float32)
            p = np.exp2(r * scale_log2 + neg_m[:, None]).astype(np.float16).astype(np.float32)   # P is fp16 in TMEM
            l[h] += p.sum(1)            # tensor-core row sum: of the fp16 P
            O[h] += p @ V[c0:c0 + 32]
    m_all = np.maximum(m_run[0], m_run[1])
    w = np.exp2(m_run - m_all[None])
    inv = 1.0 / (w[0] * l[0] + w[1] * l[1])
    return (w[0][:, None] * O[0] + w[1][:, None] * O[1]) * inv[:, None]


rng = np.random.default_rng(0)
worst = 0.0
for kv_len, gain in [(1000, 1.0), (1000, 6.0), (77, 8.0), (20, 1.0), (33, 4.0), (64, 1.0), (4096, 3.0)]:
    d, rows = 40, 64
    Lp = (kv_len + 63) // 64 * 64
    q = rng.standard_normal((rows, d)).astype(np.float16).astype(np.float32) * gain
    k = np.zeros((Lp, d), np.float32)
    k[:kv_len] = rng.standard_normal((kv_len, d)).astype(np.float16)
    v = np.zeros((Lp, d), np.float32)
    v[:kv_len] = rng.standard_normal((kv_len, d)).astype(np.float16)
    S = q @ k.T
    scale = 1 / np.sqrt(d)
    out = wide_row_block(S, v, kv_len, np.float32(scale * 1.4426950408889634))
    s = S[:, :kv_len] * scale
    p = np.exp(s - s.max(1, keepdims=True))
    ref = (p / p.sum(1, keepdims=True)) @ v[:kv_len]
    err = np.abs(out - ref).max()
    worst = max(worst, err / max(1.0, np.abs(ref).max()))
    print(f"kv_len={kv_len} gain={gain}: max abs err {err:.2e} (ref max {np.abs(ref).max():.2f}) finite={np.isfinite(out).all()}")
print("worst err / max(1, |ref|):", worst, "OK" if worst < 2e-3 else "FAIL")
