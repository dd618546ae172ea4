// FRESCO attention forward (spatial-guided and cross-frame SDPA) for sm_100a -- v2 "dual-stream".
//
// Replaces the two dense F.scaled_dot_product_attention calls of the reference processor
// (src/diffusion_hacked.py:281-285 and :303-305).
//
// One CTA owns a 128-row query tile of one (batch, head).  K/V are streamed in 64-row tiles; the
// even tiles feed softmax stream A (warps 0-3), the odd tiles stream B (warps 4-7).  Each stream
// is a complete online-softmax pipeline of its own (own running max / sum / output accumulator,
// own S, P and O regions in TMEM, own mbarriers), so the two streams never synchronise per tile;
// their partial results are merged once at the end (split-KV combine).  Four independent streams
// per SM (2 CTAs) keep the MUFU and the tensor pipe busy while any one stream waits for its MMAs.
//
//   warp 8   TMA producer   Q once, K/V tiles through a kStages-deep mbarrier ring
//   warp 9   MMA issuer     S_s = Q K_t^T  (tcgen05.mma SS, M128 N64, fp32 in TMEM)
//                           O_s = P_s V_t  (tcgen05.mma TS, P read from TMEM, V MN-major)
//   warps 0-7 softmax       one query row per thread (= one TMEM lane): the 64 scores of the tile are
//                           read from TMEM once into registers; row max, p = exp2(s*scale*log2e - m)
//                           with packed fp32x2 math, P written back to TMEM as fp16 over the first
//                           half of S.  O accumulates in TMEM across tiles; the running max is only
//                           raised (and O rescaled in TMEM) when it grows by more than 2^8, so the
//                           common tile costs no O traffic at all.
//
// Token-major [batch, tokens, heads*head_dim] fp16 tensors are consumed in place: the TMA tensor
// map views them as {head_dim, heads, tokens, batch}; a {64,1,rows,1} box lands one head's tile
// in the canonical 128B-swizzled K-major layout; columns >= head_dim and rows >= tokens are
// hardware zero-filled.
#include "common.cuh"
#include "fresco_internal.h"

namespace fresco {

constexpr int kTileM = 128;            // query rows per CTA
constexpr int kTileN = 64;             // kv rows per tile (one stream step)
constexpr int kQAtomBytes = 128 * 128;  // [128 rows x 64 fp16]
constexpr int kKVAtomBytes = 64 * 128;  // [ 64 rows x 64 fp16]
constexpr int kSoftmaxThreads = 256;
constexpr int kThreads = 320;

template <int D>
struct AttnCfg {
  static constexpr int NATOM = (D + 63) / 64;
  static constexpr int KSTEPS = (D + 15) / 16;
  static constexpr int DPAD = KSTEPS * 16;
  static constexpr int N0 = DPAD < 64 ? DPAD : 64;   // PV columns from atom 0
  static constexpr int N1 = DPAD - N0;               // PV columns from atom 1
  static constexpr bool SMALL = DPAD <= 64;
  static constexpr int TMEM_COLS = SMALL ? 256 : 512;
  static constexpr int STREAM_STRIDE = SMALL ? 128 : 256;   // TMEM columns between the two streams
  static constexpr int O_OFF = 64;                          // O region inside a stream (after S/P)
  // even, so that stream s only ever touches ring stages of parity s: the two streams then never wait on each
  // other's tiles (an odd depth can deadlock the single MMA thread)
  static constexpr int STAGES = 4;
  static constexpr int Q_BYTES = NATOM * kQAtomBytes;
  static constexpr int STAGE_BYTES = 2 * NATOM * kKVAtomBytes;
  static constexpr int SMEM_BYTES = 1024 + Q_BYTES + STAGES * STAGE_BYTES + 256;
  static constexpr int MIN_CTAS = (SMALL && SMEM_BYTES <= 110 * 1024) ? 2 : 1;
};

struct AttnParams {
  __half* out;
  int q_len, kv_len, heads, q_per_kv;
  float scale_log2;        // softmax_scale * log2(e)
  float diag_bias_log2;    // bias added where kv index == query index, * log2(e)
};

__device__ __forceinline__ unsigned long long pack_f2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

__device__ __forceinline__ void tmem_ld16_sync(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8_sync(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3])
               : "memory");
}
// wait for the outstanding tcgen05.ld's; the registers are listed as in/out operands so that no use of
// them can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait_dep64(uint32_t (&r)[64]) {
#define FR8(b) "+r"(r[b]), "+r"(r[b + 1]), "+r"(r[b + 2]), "+r"(r[b + 3]), "+r"(r[b + 4]), "+r"(r[b + 5]), "+r"(r[b + 6]), "+r"(r[b + 7])
  asm volatile("tcgen05.wait::ld.sync.aligned;" : FR8(0), FR8(8), FR8(16), FR8(24) : : "memory");
  asm volatile("" : FR8(32), FR8(40), FR8(48), FR8(56) : : "memory");
#undef FR8
}

template <int D>
__global__ void __launch_bounds__(kThreads, AttnCfg<D>::MIN_CTAS)
fresco_attn_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const AttnParams p) {
  using Cfg = AttnCfg<D>;
  constexpr int ST = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_q = smem;
  uint8_t* s_kv = smem + Cfg::Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::Q_BYTES + ST * Cfg::STAGE_BYTES);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_kv_full = bars + 1;            // [ST]
  uint64_t* bar_kv_empty = bars + 1 + ST;      // [ST]
  uint64_t* bar_s = bars + 1 + 2 * ST;         // [2]  S_s ready
  uint64_t* bar_p = bar_s + 2;                 // [2]  P_s written (128 arrivals)
  uint64_t* bar_o = bar_s + 4;                 // [2]  O_s = P_s V ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTileM;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int b_kv = b / p.q_per_kv;
  const int n_tiles = (p.kv_len + kTileN - 1) / kTileN;

  if (warp == 9 && lane == 0) {
    mbar_init(bar_q, 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(bar_kv_full + s, 1);
      mbar_init(bar_kv_empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_s + s, 1);
      mbar_init(bar_p + s, 128);
      mbar_init(bar_o + s, 1);
    }
    fence_barrier_init();
  }
  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(&tm_k);
      tma_prefetch_desc(&tm_v);
    }
    __syncwarp();
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(bar_q, Cfg::Q_BYTES);
      for (int a = 0; a < Cfg::NATOM; ++a) tma_load_4d(s_q + a * kQAtomBytes, &tm_q, bar_q, a * 64, head, q0, b);
      for (int t = 0; t < n_tiles; ++t) {
        const int st = t % ST;
        if (t >= ST) {
          const uint32_t ph = ((t / ST) - 1) & 1;
          while (!mbar_try_wait(bar_kv_empty + st, ph)) __nanosleep(64);
        }
        uint8_t* sk = s_kv + st * Cfg::STAGE_BYTES;
        uint8_t* sv = sk + Cfg::NATOM * kKVAtomBytes;
        mbar_expect_tx(bar_kv_full + st, Cfg::STAGE_BYTES);
        for (int a = 0; a < Cfg::NATOM; ++a) {
          tma_load_4d(sk + a * kKVAtomBytes, &tm_k, bar_kv_full + st, a * 64, head, t * kTileN, b_kv);
          tma_load_4d(sv + a * kKVAtomBytes, &tm_v, bar_kv_full + st, a * 64, head, t * kTileN, b_kv);
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------ MMA issuer (serves both streams)
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_f16(kTileM, kTileN, 0);
      constexpr uint32_t idesc_pv0 = make_idesc_f16(kTileM, Cfg::N0, 1);
      constexpr uint32_t idesc_pv1 = make_idesc_f16(kTileM, Cfg::N1 > 0 ? Cfg::N1 : 16, 1);
      const uint32_t q_addr = smem_u32(s_q);
      // Non-blocking state machine over the two streams: the single MMA thread never parks on one stream's
      // barrier while the other stream has work (and the K/V ring can therefore never deadlock).
      //   state 0: next tile's K/V not yet confirmed in smem -> poll kv_full, then issue S = Q K^T
      //   state 1: S issued, waiting for the stream's P          -> poll bar_p,  then issue O = P V
      auto issue_qk = [&](int s, int t) {
        const int st = t % ST;
        tc_fence_after();
        const uint32_t k_addr = smem_u32(s_kv + st * Cfg::STAGE_BYTES);
        const uint32_t d_tmem = tmem + s * Cfg::STREAM_STRIDE;
#pragma unroll
        for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
          const uint32_t qoff = (ks >> 2) * kQAtomBytes + (ks & 3) * 32;
          const uint32_t koff = (ks >> 2) * kKVAtomBytes + (ks & 3) * 32;
          umma_ss(d_tmem, make_smem_desc_sw128(q_addr + qoff, 16, 1024), make_smem_desc_sw128(k_addr + koff, 16, 1024),
                  idesc_qk, ks > 0);
        }
        umma_commit(bar_s + s);
      };
      auto issue_pv = [&](int s, int t, bool first, bool last) {
        const int st = t % ST;
        tc_fence_after();
        const uint32_t v_addr = smem_u32(s_kv + st * Cfg::STAGE_BYTES + Cfg::NATOM * kKVAtomBytes);
        const uint32_t p_tmem = tmem + s * Cfg::STREAM_STRIDE;
        const uint32_t o_tmem = p_tmem + Cfg::O_OFF;
#pragma unroll
        for (int k2 = 0; k2 < kTileN / 16; ++k2) {
          const uint32_t acc = (k2 > 0 || !first) ? 1u : 0u;      // O accumulates in TMEM across the stream's tiles
          umma_ts(o_tmem, p_tmem + k2 * 8, make_smem_desc_sw128(v_addr + k2 * 2048, kKVAtomBytes, 1024), idesc_pv0,
                  acc);
          if (Cfg::N1 > 0)
            umma_ts(o_tmem + 64, p_tmem + k2 * 8,
                    make_smem_desc_sw128(v_addr + kKVAtomBytes + k2 * 2048, kKVAtomBytes, 1024), idesc_pv1, acc);
        }
        umma_commit(bar_kv_empty + st);
        if (last) umma_commit(bar_o + s);                         // single-phase "stream finished" signal
      };
      mbar_wait(bar_q, 0, 1);
      int it[2] = {0, 0};                     // per-stream step counter; stream s handles tiles s, s+2, ...
      int state[2] = {0, 0};
      const int cnt[2] = {(n_tiles + 1) / 2, n_tiles / 2};
      int remaining = cnt[0] + cnt[1];
      uint32_t idle = 0;
      while (remaining > 0) {
        bool progressed = false;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (it[s] >= cnt[s]) continue;
          const int t = s + 2 * it[s];
          if (state[s] == 0) {
            if (mbar_try_wait(bar_kv_full + (t % ST), (t / ST) & 1)) {
              issue_qk(s, t);
              state[s] = 1;
              progressed = true;
            }
          } else if (mbar_try_wait(bar_p + s, it[s] & 1)) {
            issue_pv(s, t, it[s] == 0, it[s] + 1 == cnt[s]);
            ++it[s];
            --remaining;
            state[s] = 0;
            progressed = true;
          }
        }
        if (progressed) {
          idle = 0;
        } else if (++idle > FRESCO_WATCHDOG_POLLS) {
          printf("fresco_b200 watchdog: attention MMA issuer stalled, block (%d,%d,%d) it=(%d,%d)/(%d,%d) state=(%d,%d)\n",
                 blockIdx.x, blockIdx.y, blockIdx.z, it[0], it[1], cnt[0], cnt[1], state[0], state[1]);
          __trap();
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax streams
    const int s = warp >> 2;                               // stream: 0 = even tiles, 1 = odd tiles
    const int row = (warp & 3) * 32 + lane;                // query row inside the tile == TMEM lane
    const uint32_t t_lane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + s * Cfg::STREAM_STRIDE;
    const int q_row = q0 + row;
    const int my_tiles = s == 0 ? (n_tiles + 1) / 2 : n_tiles / 2;
    const int kv_len = p.kv_len;
    const float scale_log2 = p.scale_log2, bias_log2 = p.diag_bias_log2;
    const bool use_bias = bias_log2 != 0.f;
    const unsigned long long scale2 = pack_f2(scale_log2, scale_log2);
    float m_run = -INFINITY, l_run = 0.f;

    for (int i = 0; i < my_tiles; ++i) {
      const int col0 = (s + 2 * i) * kTileN;
      // warp-uniform: does this tile need masking (ragged tail) or the diagonal bias?
      const bool special = (col0 + kTileN > kv_len) ||
                           (use_bias && (q0 + (warp & 3) * 32) < col0 + kTileN && (q0 + (warp & 3) * 32 + 32) > col0);
      mbar_wait(bar_s + s, i & 1, 2);          // S_i ready; this also implies the stream's previous P V has retired
      tc_fence_after();
      // ---- the whole 64-column row of scores, once, into registers
      uint32_t r[64];
      tmem_ld16(t_lane + 0, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
      tmem_ld16(t_lane + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
      tmem_ld16(t_lane + 32, *reinterpret_cast<uint32_t(*)[16]>(&r[32]));
      tmem_ld16(t_lane + 48, *reinterpret_cast<uint32_t(*)[16]>(&r[48]));
      tmem_ld_wait_dep64(r);
      if (special) {                            // rare path: fold mask / bias into the raw scores
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          const int col = col0 + j;
          float v = __uint_as_float(r[j]);
          if (use_bias && col == q_row) v += bias_log2 / scale_log2;
          if (col >= kv_len) v = -INFINITY;
          r[j] = __float_as_uint(v);
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 64; j += 8) {
        mx0 = max3(mx0, __uint_as_float(r[j]), __uint_as_float(r[j + 1]));
        mx1 = max3(mx1, __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        mx2 = max3(mx2, __uint_as_float(r[j + 4]), __uint_as_float(r[j + 5]));
        mx3 = max3(mx3, __uint_as_float(r[j + 6]), __uint_as_float(r[j + 7]));
      }
      const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
      // ---- lazy running max: raise it (and rescale O in TMEM) only when it grows by more than 2^8
      if (i == 0) {
        m_run = m_tile;
      } else {
        const bool need = m_tile > m_run + 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? fast_exp2(m_run - m_tile) : 1.0f;
          if (need) {
            l_run *= alpha;
            m_run = m_tile;
          }
#pragma unroll
          for (int c = 0; c < D / 8; ++c) {
            uint32_t o[8];
            const int col = c * 8;
            const uint32_t addr = t_lane + Cfg::O_OFF + (col < Cfg::N0 ? col : 64 + (col - Cfg::N0));
            tmem_ld8_sync(addr, o);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
            tmem_st8(addr, o);
          }
        }
      }
      // ---- p = exp2(s*scale - m), packed to fp16 over the first 32 columns of S
      const unsigned long long negm2 = pack_f2(-m_run, -m_run);
      unsigned long long sum2a = pack_f2(0.f, 0.f), sum2b = pack_f2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          float t0, t1;
          unpack_f2(fma2(pack_f2(__uint_as_float(r[c * 8 + j]), __uint_as_float(r[c * 8 + j + 1])), scale2, negm2), t0, t1);
          const float p0 = fast_exp2(t0);
          const float p1 = fast_exp2(t1);
          if (j & 2) sum2b = add2(sum2b, pack_f2(p0, p1)); else sum2a = add2(sum2a, pack_f2(p0, p1));
          pk[j >> 1] = pack_half2(p0, p1);
        }
        tmem_st4(t_lane + c * 4, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_p + s);
      float sa, sb;
      unpack_f2(add2(sum2a, sum2b), sa, sb);
      l_run += sa + sb;
    }

    // ---- stream epilogue: fetch the accumulated O from TMEM
    float o_acc[D];
    if (my_tiles > 0) {
      mbar_wait(bar_o + s, 0, 3);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < D / 8; ++c) {
        uint32_t o[8];
        const int col = c * 8;
        tmem_ld8_sync(t_lane + Cfg::O_OFF + (col < Cfg::N0 ? col : 64 + (col - Cfg::N0)), o);
#pragma unroll
        for (int j = 0; j < 8; ++j) o_acc[col + j] = __uint_as_float(o[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < D; ++j) o_acc[j] = 0.f;
    }

    // ---- merge the two streams (split-KV combine) and store
    // all MMAs that read the K/V ring have completed once both streams saw their last bar_o, so the ring is free
    float* xch = reinterpret_cast<float*>(s_kv);                       // [D + 2][128]
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (s == 1) {
      xch[0 * 128 + row] = m_run;
      xch[1 * 128 + row] = l_run;
#pragma unroll
      for (int i = 0; i < D; ++i) xch[(2 + i) * 128 + row] = o_acc[i];
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (s == 0 && q_row < p.q_len) {
      const float mb = xch[row], lb = xch[128 + row];
      const float m = fmaxf(m_run, mb);
      const float wa = fast_exp2(m_run - m);
      const float wb = (mb == -INFINITY) ? 0.f : fast_exp2(mb - m);
      const float inv = 1.f / (l_run * wa + lb * wb);
      const float ca = wa * inv, cb = wb * inv;
      __half* dst = p.out + (static_cast<size_t>(b) * p.q_len + q_row) * (static_cast<size_t>(p.heads) * D) +
                    static_cast<size_t>(head) * D;
#pragma unroll
      for (int v8 = 0; v8 < D / 8; ++v8) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = o_acc[v8 * 8 + j] * ca + xch[(2 + v8 * 8 + j) * 128 + row] * cb;
        uint4 pkt;
        pkt.x = pack_half2(o[0], o[1]);
        pkt.y = pack_half2(o[2], o[3]);
        pkt.z = pack_half2(o[4], o[5]);
        pkt.w = pack_half2(o[6], o[7]);
        reinterpret_cast<uint4*>(dst)[v8] = pkt;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

// {head_dim, heads, tokens, batch} view of a token-major [batch, tokens, heads*head_dim] fp16 tensor
static int make_head_tile_map(CUtensorMap* map, const void* base, int head_dim, int heads, int tokens, int batch,
                              int box_rows) {
  const cuuint64_t dims[4] = {(cuuint64_t)head_dim, (cuuint64_t)heads, (cuuint64_t)tokens, (cuuint64_t)batch};
  const cuuint64_t strides[3] = {(cuuint64_t)head_dim * 2, (cuuint64_t)heads * head_dim * 2,
                                 (cuuint64_t)tokens * heads * head_dim * 2};
  const cuuint32_t box[4] = {64, 1, (cuuint32_t)box_rows, 1};
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
  if (make_head_tile_map(&tq, q, D, heads, q_len, batch_q, kTileM)) return FRESCO_ERR_TENSORMAP;
  if (make_head_tile_map(&tk, k, D, heads, kv_len, batch_kv, kTileN)) return FRESCO_ERR_TENSORMAP;
  if (make_head_tile_map(&tv, v, D, heads, kv_len, batch_kv, kTileN)) return FRESCO_ERR_TENSORMAP;
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
