// FRESCO attention forward (spatial-guided and cross-frame SDPA) for sm_100a.
//
// Replaces the two dense F.scaled_dot_product_attention calls of the reference
// processor (src/diffusion_hacked.py:281-285 and :303-305).  One CTA owns a
// 128-row query tile of one (batch, head) and streams 128-row K/V tiles:
//
//   warp 4   TMA producer      Q once, K/V tiles through a 2-stage mbarrier ring
//   warp 5   MMA issuer        S = Q K^T   (tcgen05.mma SS, fp32 accum in TMEM)
//                              PV = P V    (tcgen05.mma TS, P read from TMEM)
//   warps 0-3 softmax          one query row per thread (= one TMEM lane):
//                              online softmax, P written back to TMEM as fp16
//                              (aliasing S), running O kept in registers
//
// Token-major [batch, tokens, heads*head_dim] fp16 tensors are consumed in
// place: the TMA tensor map views them as {head_dim, heads, tokens, batch} and a
// {64,1,128,1} box lands one head's [128 x 64] tile in the canonical 128B-swizzled
// K-major layout; columns >= head_dim and rows >= tokens are hardware zero-filled.
#include "common.cuh"
#include "fresco_internal.h"

namespace fresco {

constexpr int kTileM = 128;         // query rows per CTA
constexpr int kTileN = 128;         // kv rows per step
constexpr int kAtomBytes = 128 * 128;  // [128 rows x 64 fp16] swizzle atom
constexpr int kStages = 2;
constexpr int kSoftmaxThreads = 128;
constexpr int kThreads = 192;
constexpr int kTmemCols = 256;
constexpr int kTmemO = 128;         // column offset of the PV accumulator

template <int D>
struct AttnCfg {
  static constexpr int NATOM = (D + 63) / 64;
  static constexpr int KSTEPS = (D + 15) / 16;
  static constexpr int DPAD = KSTEPS * 16;
  static constexpr int N0 = DPAD < 64 ? DPAD : 64;   // PV columns from atom 0
  static constexpr int N1 = DPAD - N0;               // PV columns from atom 1
  static constexpr int Q_BYTES = NATOM * kAtomBytes;
  static constexpr int STAGE_BYTES = 2 * NATOM * kAtomBytes;
  static constexpr int SMEM_BYTES = 1024 + Q_BYTES + kStages * STAGE_BYTES + 256;
};

struct AttnParams {
  __half* out;
  int q_len, kv_len, heads, q_per_kv;
  float scale_log2;        // softmax_scale * log2(e)
  float diag_bias_log2;    // bias added where kv index == query index, * log2(e)
};

template <int D>
__global__ void __launch_bounds__(kThreads, (AttnCfg<D>::SMEM_BYTES <= 110 * 1024) ? 2 : 1)
fresco_attn_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const AttnParams p) {
  using Cfg = AttnCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_q = smem;
  uint8_t* s_kv = smem + Cfg::Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::Q_BYTES + kStages * Cfg::STAGE_BYTES);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_kv_full = bars + 1;              // [kStages]
  uint64_t* bar_kv_empty = bars + 1 + kStages;   // [kStages]
  uint64_t* bar_s = bars + 1 + 2 * kStages;
  uint64_t* bar_p = bar_s + 1;
  uint64_t* bar_o = bar_s + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 3);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTileM;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int b_kv = b / p.q_per_kv;
  const int n_tiles = (p.kv_len + kTileN - 1) / kTileN;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  if (warp == 5 && lane == 0) {
    mbar_init(bar_q, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_kv_full + s, 1);
      mbar_init(bar_kv_empty + s, 1);
    }
    mbar_init(bar_s, 1);
    mbar_init(bar_p, kSoftmaxThreads);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(bar_q, Cfg::Q_BYTES);
      for (int a = 0; a < Cfg::NATOM; ++a) tma_load_4d(s_q + a * kAtomBytes, &tm_q, bar_q, a * 64, head, q0, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % kStages;
        if (j >= kStages) mbar_wait(bar_kv_empty + st, ((j / kStages) - 1) & 1);
        uint8_t* sk = s_kv + st * Cfg::STAGE_BYTES;
        uint8_t* sv = sk + Cfg::NATOM * kAtomBytes;
        mbar_expect_tx(bar_kv_full + st, Cfg::STAGE_BYTES);
        for (int a = 0; a < Cfg::NATOM; ++a) {
          tma_load_4d(sk + a * kAtomBytes, &tm_k, bar_kv_full + st, a * 64, head, j * kTileN, b_kv);
          tma_load_4d(sv + a * kAtomBytes, &tm_v, bar_kv_full + st, a * 64, head, j * kTileN, b_kv);
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_f16(kTileM, kTileN, 0);
      constexpr uint32_t idesc_pv0 = make_idesc_f16(kTileM, Cfg::N0, 1);
      constexpr uint32_t idesc_pv1 = make_idesc_f16(kTileM, Cfg::N1 > 0 ? Cfg::N1 : 16, 1);
      const uint32_t q_addr = smem_u32(s_q);
      auto issue_qk = [&](int st) {
        const uint32_t k_addr = smem_u32(s_kv + st * Cfg::STAGE_BYTES);
#pragma unroll
        for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
          const uint32_t off = (ks >> 2) * kAtomBytes + (ks & 3) * 32;   // 16 fp16 = 32 B inside the swizzle row
          umma_ss(tmem, make_smem_desc_sw128(q_addr + off, 16, 1024), make_smem_desc_sw128(k_addr + off, 16, 1024),
                  idesc_qk, ks > 0);
        }
      };
      auto issue_pv = [&](int st) {
        const uint32_t v_addr = smem_u32(s_kv + st * Cfg::STAGE_BYTES + Cfg::NATOM * kAtomBytes);
#pragma unroll
        for (int k2 = 0; k2 < kTileN / 16; ++k2) {
          // P: 16 fp16 along K = 8 TMEM columns; V: 16 kv rows = 2 swizzle row-groups = 2048 B
          umma_ts(tmem + kTmemO, tmem + k2 * 8, make_smem_desc_sw128(v_addr + k2 * 2048, kAtomBytes, 1024),
                  idesc_pv0, k2 > 0);
          if (Cfg::N1 > 0)
            umma_ts(tmem + kTmemO + 64, tmem + k2 * 8,
                    make_smem_desc_sw128(v_addr + kAtomBytes + k2 * 2048, kAtomBytes, 1024), idesc_pv1, k2 > 0);
        }
      };
      mbar_wait(bar_q, 0);
      mbar_wait(bar_kv_full + 0, 0);
      tc_fence_after();
      issue_qk(0);
      umma_commit(bar_s);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % kStages;
        mbar_wait(bar_p, j & 1);
        tc_fence_after();
        issue_pv(st);
        umma_commit(bar_kv_empty + st);
        umma_commit(bar_o);
        if (j + 1 < n_tiles) {
          const int st2 = (j + 1) % kStages;
          mbar_wait(bar_kv_full + st2, ((j + 1) / kStages) & 1);
          tc_fence_after();
          issue_qk(st2);
          umma_commit(bar_s);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax / output warps
    const int row = threadIdx.x;                         // query row inside the tile == TMEM lane
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const int q_row = q0 + row;
    float m_run = -INFINITY, l_run = 0.f;
    float o_acc[Cfg::DPAD];
#pragma unroll
    for (int i = 0; i < Cfg::DPAD; ++i) o_acc[i] = 0.f;
    const bool use_bias = p.diag_bias_log2 != 0.f;

    for (int j = 0; j < n_tiles; ++j) {
      const int col0 = j * kTileN;
      const bool tail = col0 + kTileN > p.kv_len;
      const bool special = tail || (use_bias && q_row >= col0 && q_row < col0 + kTileN);
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      // pass 1: row maximum of t = s * scale_log2 (+ bias on the diagonal)
      float m_tile = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(t_lane + c * 32, r);
        tmem_ld_wait();
        if (!special) {
#pragma unroll
          for (int i = 0; i < 32; ++i) m_tile = fmaxf(m_tile, __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int col = col0 + c * 32 + i;
            float t = __uint_as_float(r[i]);
            if (use_bias && col == q_row) t += p.diag_bias_log2 / p.scale_log2;
            if (col >= p.kv_len) t = -INFINITY;
            m_tile = fmaxf(m_tile, t);
          }
        }
      }
      m_tile *= p.scale_log2;
      const float m_new = fmaxf(m_run, m_tile);
      const float alpha = fast_exp2(m_run - m_new);
      float rowsum = 0.f;
      // pass 2: p = exp2(t - m_new), packed to fp16 and written over S
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(t_lane + c * 32, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float t0 = fmaf(__uint_as_float(r[i]), p.scale_log2, -m_new);
          float t1 = fmaf(__uint_as_float(r[i + 1]), p.scale_log2, -m_new);
          if (special) {
            const int col = col0 + c * 32 + i;
            if (use_bias && col == q_row) t0 += p.diag_bias_log2;
            if (use_bias && col + 1 == q_row) t1 += p.diag_bias_log2;
            if (col >= p.kv_len) t0 = -INFINITY;
            if (col + 1 >= p.kv_len) t1 = -INFINITY;
          }
          const float p0 = fast_exp2(t0);
          const float p1 = fast_exp2(t1);
          rowsum += p0 + p1;
          pk[i >> 1] = pack_half2(p0, p1);
        }
        tmem_st16(t_lane + c * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_p);
      l_run = l_run * alpha + rowsum;
      m_run = m_new;
      // fold this tile's P V into the running output
      mbar_wait(bar_o, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < Cfg::DPAD / 16; ++c) {
        uint32_t r[16];
        tmem_ld16(t_lane + kTmemO + (c < Cfg::N0 / 16 ? c * 16 : 64 + (c - Cfg::N0 / 16) * 16), r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o_acc[c * 16 + i] = fmaf(o_acc[c * 16 + i], alpha, __uint_as_float(r[i]));
      }
      tc_fence_before();
    }
    // epilogue: normalise and store this row's head slice (D fp16, 16-byte vectors)
    if (q_row < p.q_len) {
      const float inv = 1.f / l_run;
      __half* dst = p.out + (static_cast<size_t>(b) * p.q_len + q_row) * (static_cast<size_t>(p.heads) * D) +
                    static_cast<size_t>(head) * D;
#pragma unroll
      for (int v8 = 0; v8 < D / 8; ++v8) {
        uint4 pkt;
        pkt.x = pack_half2(o_acc[v8 * 8 + 0] * inv, o_acc[v8 * 8 + 1] * inv);
        pkt.y = pack_half2(o_acc[v8 * 8 + 2] * inv, o_acc[v8 * 8 + 3] * inv);
        pkt.z = pack_half2(o_acc[v8 * 8 + 4] * inv, o_acc[v8 * 8 + 5] * inv);
        pkt.w = pack_half2(o_acc[v8 * 8 + 6] * inv, o_acc[v8 * 8 + 7] * inv);
        reinterpret_cast<uint4*>(dst)[v8] = pkt;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<kTmemCols>(tmem);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

// {head_dim, heads, tokens, batch} view of a token-major [batch, tokens, heads*head_dim] fp16 tensor
static int make_head_tile_map(CUtensorMap* map, const void* base, int head_dim, int heads, int tokens, int batch) {
  const cuuint64_t dims[4] = {(cuuint64_t)head_dim, (cuuint64_t)heads, (cuuint64_t)tokens, (cuuint64_t)batch};
  const cuuint64_t strides[3] = {(cuuint64_t)head_dim * 2, (cuuint64_t)heads * head_dim * 2,
                                 (cuuint64_t)tokens * heads * head_dim * 2};
  const cuuint32_t box[4] = {64, 1, 128, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return encode_tiled_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int D>
static int launch_attn(const void* q, const void* k, const void* v, void* out, int batch_q, int q_len, int kv_len,
                       int heads, int q_per_kv, float softmax_scale, float diag_bias, cudaStream_t stream) {
  using Cfg = AttnCfg<D>;
  CUtensorMap tq, tk, tv;
  const int batch_kv = batch_q / q_per_kv;
  if (make_head_tile_map(&tq, q, D, heads, q_len, batch_q)) return FRESCO_ERR_TENSORMAP;
  if (make_head_tile_map(&tk, k, D, heads, kv_len, batch_kv)) return FRESCO_ERR_TENSORMAP;
  if (make_head_tile_map(&tv, v, D, heads, kv_len, batch_kv)) return FRESCO_ERR_TENSORMAP;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fresco_attn_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(attn)");
    attr_set = true;
  }
  AttnParams p;
  p.out = static_cast<__half*>(out);
  p.q_len = q_len;
  p.kv_len = kv_len;
  p.heads = heads;
  p.q_per_kv = q_per_kv;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.diag_bias_log2 = diag_bias * 1.4426950408889634f;
  dim3 grid((q_len + kTileM - 1) / kTileM, heads, batch_q);
  fresco_attn_kernel<D><<<grid, kThreads, Cfg::SMEM_BYTES, stream>>>(tq, tk, tv, p);
  return check_launch("fresco_attn_kernel");
}

}  // namespace fresco

using namespace fresco;

extern "C" int fresco_attn_fwd(const void* q, const void* k, const void* v, void* out, int batch_q, int q_len,
                               int kv_len, int heads, int head_dim, int q_per_kv, float softmax_scale,
                               float diag_bias, void* stream) {
  if (!q || !k || !v || !out) return set_error(FRESCO_ERR_ARG, "fresco_attn_fwd: null pointer");
  if (batch_q <= 0 || q_len <= 0 || kv_len <= 0 || heads <= 0 || q_per_kv <= 0 || batch_q % q_per_kv != 0)
    return set_error(FRESCO_ERR_ARG, "fresco_attn_fwd: bad shape");
  if (softmax_scale <= 0.f) return set_error(FRESCO_ERR_ARG, "fresco_attn_fwd: softmax_scale must be > 0");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (head_dim) {
    case 40: return launch_attn<40>(q, k, v, out, batch_q, q_len, kv_len, heads, q_per_kv, softmax_scale, diag_bias, s);
    case 64: return launch_attn<64>(q, k, v, out, batch_q, q_len, kv_len, heads, q_per_kv, softmax_scale, diag_bias, s);
    case 80: return launch_attn<80>(q, k, v, out, batch_q, q_len, kv_len, heads, q_per_kv, softmax_scale, diag_bias, s);
    case 128: return launch_attn<128>(q, k, v, out, batch_q, q_len, kv_len, heads, q_per_kv, softmax_scale, diag_bias, s);
    default: return set_error(FRESCO_ERR_UNSUPPORTED, "fresco_attn_fwd: head_dim must be one of 40, 64, 80, 128");
  }
}
