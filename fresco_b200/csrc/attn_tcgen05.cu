// FRESCO attention forward (spatial-guided and cross-frame SDPA) for sm_100a.
//
// Replaces the two dense F.scaled_dot_product_attention calls of the reference processor
// (src/diffusion_hacked.py:281-285 and :303-305).
//
// Common to both kernels in this file: tcgen05 / TMEM / TMA flash attention.  Token-major [batch, tokens, heads*head_dim]
// fp16 tensors are consumed in place -- the TMA tensor map views them as {head_dim, heads, tokens, batch}; a
// {64,1,rows,1} box lands one head's tile in the canonical 128B-swizzled K-major layout; columns >= head_dim and rows
// >= tokens are hardware zero-filled (no padding passes for head_dim 40 / 80 or ragged lengths).  S = Q K^T by
// tcgen05.mma kind::f16 SS into TMEM (fp32); softmax threads own TMEM lanes (= query rows): tcgen05.ld, row max
// (FMNMX3), p = exp2(s*scale*log2e - m) with packed fma.rn.f32x2 and, for every n-th pair, a polynomial on the FMA
// pipe; P goes back to TMEM as fp16 (tcgen05.st) and O += P V runs as a TS MMA (A = P from TMEM, B = V read MN-major
// from the same swizzled tile), accumulated in TMEM over all key tiles; the running max is lazy (O is rescaled only
// when a tile exceeds it by more than 2^8); row sums come out of the tensor core.
//
//   fresco_attn_twin_kernel   (head_dim <= 80: the FRESCO shapes) two 128-row query tiles per CTA, 128-key tiles, one CTA
//                             per SM; see the comment above the kernel.
//   fresco_attn_kernel        ("pipelined", every head_dim; the default above 80: GMFlow's d = 128) one 128-row query
//                             tile per CTA, 64-key tiles, S and P double-buffered, two CTAs per SM at head_dim <= 64:
//
//   warps 0-3   softmax       one query row per thread.  The 64 scores of a tile are read from TMEM once into registers.
//   producer    TMA           Q once, K/V tiles through a 4/5-stage mbarrier ring
//   QK issuer   tcgen05.mma   S_i = Q K_i^T (M128 N64) into one of TWO S buffers, issued as soon as the softmax threads
//                             hold S_{i-2} in registers, so scores are always ready ahead of time
//   PV issuer   tcgen05.mma   O += P_i V_i; for head_dim 40 one more N=16 MMA per K-step against a tile of ones yields
//                             the row sums
//   With one CTA per SM (head_dim 80/128) each of the three roles has its own warp (224 threads).  With two CTAs per
//   SM (head_dim 40/64) the producer shares a thread with the QK issuer and refills the ring without ever blocking.
//   TMEM: S0 64 + S1 64 + P0 32 + P1 32 + O <= 64 columns = 256 -> two CTAs per SM for d <= 64.
//
// What bounds them (measured: tools/pipe_probe.cu, tools/softmax_mix_probe.cu, profiles/r02_attn_*_hot*.txt): at
// head_dim 40 a 128 x 64 tile is 192 tensor-core clocks but 512 MUFU clocks, and every 64-key structure that was tried
// ended at ~840 clocks per tile with the issue slots 56-68 % busy: 368 (pipelined) to 572 warp-instructions per tile, of
// which 161 are the softmax arithmetic.  The twin kernel pays the per-tile overhead once per 128 keys and has no
// P-buffer / O-stability waits by construction; its exponentials are split between the MUFU and the FMA pipe.
#include "common.cuh"
#include "fresco_internal.h"

namespace fresco {

constexpr int kTileM = 128;            // query rows per CTA
constexpr int kTileN = 64;             // kv rows per tile
constexpr int kQAtomBytes = 128 * 128;  // [128 rows x 64 fp16]
constexpr int kKVAtomBytes = 64 * 128;  // [ 64 rows x 64 fp16]
#ifdef FRESCO_ATTN_ABLATE_BUILD
#define ABL(p, bit) ((p).ablate & (bit))
#else
#define ABL(p, bit) 0
#endif

// Profiling aid (-DFRESCO_ATTN_TRACE): SM-clock stamps of one CTA's pipeline stages for tiles [64, 96), read back with
// fresco_debug_attn_trace().  Never part of the product build.
#ifdef FRESCO_ATTN_TRACE
__device__ long long g_attn_trace[32 * 16];
#define TRACE(i, slot, dep)                                                                       \
  do {                                                                                            \
    if (trace_cta && (i) >= 64 && (i) < 96) {                                                     \
      long long t_;                                                                               \
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) : "r"(dep) : "memory");                    \
      g_attn_trace[((i) - 64) * 16 + (slot)] = t_;                                                \
    }                                                                                             \
  } while (0)
#else
#define TRACE(i, slot, dep) do { } while (0)
#endif

template <int D, bool ROWSUM = false>
struct AttnCfg {
  static constexpr int NATOM = (D + 63) / 64;
  static constexpr int KSTEPS = (D + 15) / 16;
  static constexpr int DPAD = KSTEPS * 16;
  static constexpr int N0 = DPAD < 64 ? DPAD : 64;   // PV columns from atom 0
  static constexpr int N1 = DPAD - N0;               // PV columns from atom 1
  static constexpr int S_OFF0 = 0, S_OFF1 = 64, P_OFF0 = 128, P_OFF1 = 160, O_OFF = 192;
  static constexpr int TMEM_COLS = (O_OFF + DPAD <= 256) ? 256 : 512;
  static constexpr int STAGES = NATOM == 1 ? 5 : 4;
  // Row sums l = sum_j p_j come out of the tensor core when the O region has 16 spare columns (d = 40: O uses 48 of
  // its 64): every K-step issues one extra N=16 MMA of P against a constant tile of ones into O columns [48, 64).
  // That removes the add.f32x2 chain (about 10 % of the softmax warps' instructions) from the issue-bound loop.
  static constexpr bool MMA_ROWSUM = ROWSUM && (DPAD + 16 <= 64);
  static constexpr int L_COL = 48;
  static constexpr int ONES_BYTES = MMA_ROWSUM ? 2048 : 0;
  static constexpr int Q_BYTES = NATOM * kQAtomBytes;
  static constexpr int STAGE_BYTES = 2 * NATOM * kKVAtomBytes;
  static constexpr int SMEM_BYTES = 1024 + Q_BYTES + STAGES * STAGE_BYTES + ONES_BYTES + 256;
  static constexpr int MIN_CTAS = (TMEM_COLS == 256 && SMEM_BYTES <= 112 * 1024) ? 2 : 1;
  // Warp roles.  With one CTA per SM there are registers to spare, so the TMA producer, the score-MMA issuer and the
  // P V issuer each get a warp (SPLIT).  With two CTAs per SM a seventh warp would push the softmax threads under the
  // 168 registers they need, so the producer shares a thread with the score-MMA issuer.
  static constexpr bool SPLIT = MIN_CTAS == 1;
  static constexpr int THREADS = SPLIT ? 224 : 192;
  static constexpr int QK_WARP = SPLIT ? 5 : 4, PV_WARP = SPLIT ? 6 : 5;
};

struct AttnParams {
  __half* out;
  int q_len, kv_len, heads, q_per_kv;
  float scale_log2;        // softmax_scale * log2(e)
  float diag_bias_log2;    // bias added where kv index == query index, * log2(e)
  int ablate;              // profiling aid, only honoured when built with -DFRESCO_ATTN_ABLATE_BUILD (results are
                           // WRONG when non-zero; env FRESCO_ATTN_ABLATE selects the bits): bit0 no exp2,
                           // bit1 no S load from TMEM, bit2 no P store, bit3 no P V MMA, bit4 no row max,
                           // bit5 no Q K^T MMA (commit only), bit6 K/V tiles loaded once per ring stage only
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

// exp2 on the FMA pipe for a fraction of the scores (the MUFU pipe, 16 ex2/clk/SM, is the busiest unit of this
// kernel): Cody-Waite split x = n + f, f in [-0.5, 0.5], degree-3 minimax polynomial for 2^f (max relative error
// 7.5e-5, well inside the fp16 rounding of P), exponent patched in with an integer add.  Two scores at a time so
// the range reduction and Horner steps are packed fp32x2 instructions.
__device__ __forceinline__ void exp2_poly_x2(float t0, float t1, float& p0, float& p1) {
  t0 = fmaxf(t0, -126.0f);
  t1 = fmaxf(t1, -126.0f);
  const unsigned long long t2 = pack_f2(t0, t1);
  const unsigned long long magic = pack_f2(12582912.0f, 12582912.0f);            // 1.5 * 2^23
  const unsigned long long r2 = add2(t2, magic);                                  // integer part in the low mantissa bits
  const unsigned long long n2 = add2(r2, pack_f2(-12582912.0f, -12582912.0f));
  const unsigned long long f2 = fma2(n2, pack_f2(-1.0f, -1.0f), t2);
  unsigned long long q2 = fma2(pack_f2(0.05517164617776871f, 0.05517164617776871f), f2,
                               pack_f2(0.2426111251115799f, 0.2426111251115799f));
  q2 = fma2(q2, f2, pack_f2(0.6932609677314758f, 0.6932609677314758f));
  q2 = fma2(q2, f2, pack_f2(0.9999280571937561f, 0.9999280571937561f));
  float q0, q1, r0, r1;
  unpack_f2(q2, q0, q1);
  unpack_f2(r2, r0, r1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(r0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(r1) << 23));
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

// POLY: every POLY-th pair of scores takes the FMA-pipe exp2 instead of MUFU.EX2 (0 = never)
template <int D, int POLY, bool ROWSUM>
__global__ void __launch_bounds__(AttnCfg<D>::THREADS, AttnCfg<D>::MIN_CTAS)
fresco_attn_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const AttnParams p) {
  using Cfg = AttnCfg<D, ROWSUM>;
  constexpr int ST = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_q = smem;
  uint8_t* s_kv = smem + Cfg::Q_BYTES;
  uint8_t* s_ones = smem + Cfg::Q_BYTES + ST * Cfg::STAGE_BYTES;      // [16 kv rows x 128 B] of fp16 1.0 (MMA_ROWSUM)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ones + Cfg::ONES_BYTES);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_kv_full = bars + 1;            // [ST]
  uint64_t* bar_kv_empty = bars + 1 + ST;      // [ST]
  uint64_t* bar_s = bars + 1 + 2 * ST;         // [2]  S buffer b holds tile i (i & 1 == b)
  // every barrier below exists once per TMEM buffer (index i & 1, phase (i >> 1) & 1): a parity wait is only sound
  // if the waited barrier cannot complete two phases before the waiter looks at it
  uint64_t* bar_p = bar_s + 2;                 // [2] P_i written (128 arrivals)
  uint64_t* bar_o = bar_s + 4;                 // [2] P_i V_i retired (P buffer free, O stable)
  uint64_t* bar_c = bar_s + 6;                 // [2] S_i copied to registers by all 128 softmax threads (S buffer free)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTileM;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int b_kv = b / p.q_per_kv;
  const int n_tiles = (p.kv_len + kTileN - 1) / kTileN;
#ifdef FRESCO_ATTN_TRACE
  const bool trace_cta = blockIdx.x == 3 && blockIdx.y == 2 && blockIdx.z == 0 && lane == 0 && (warp == 0 || warp >= 4);
#endif

  if (warp == 5 && lane == 0) {
    mbar_init(bar_q, 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(bar_kv_full + s, 1);
      mbar_init(bar_kv_empty + s, 2);          // released by the QK issuer (K read) and by the PV issuer (V read)
    }
    mbar_init(bar_s + 0, 1);
    mbar_init(bar_s + 1, 1);
    mbar_init(bar_p + 0, 4);                     // one elected arrival per softmax warp: 128 per-thread arrivals on one
    mbar_init(bar_p + 1, 4);                     // shared-memory word serialise (each is an atomic on the same address)
    mbar_init(bar_o + 0, 1);
    mbar_init(bar_o + 1, 1);
    mbar_init(bar_c + 0, 4);
    mbar_init(bar_c + 1, 4);
    fence_barrier_init();
  }
  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(&tm_k);
      tma_prefetch_desc(&tm_v);
    }
    __syncwarp();
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  if (Cfg::MMA_ROWSUM) {
    for (int i = threadIdx.x; i < Cfg::ONES_BYTES / 4; i += Cfg::THREADS) reinterpret_cast<uint32_t*>(s_ones)[i] = 0x3C003C00u;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // K/V tile t -> ring stage t % ST (the caller has made sure that the stage is free)
  auto load_kv_tile = [&](int t) {
    const int st = t % ST;
    uint8_t* sk = s_kv + st * Cfg::STAGE_BYTES;
    uint8_t* sv = sk + Cfg::NATOM * kKVAtomBytes;
    mbar_expect_tx(bar_kv_full + st, Cfg::STAGE_BYTES);
    for (int a = 0; a < Cfg::NATOM; ++a) {
      tma_load_4d(sk + a * kKVAtomBytes, &tm_k, bar_kv_full + st, a * 64, head, t * kTileN, b_kv);
      tma_load_4d(sv + a * kKVAtomBytes, &tm_v, bar_kv_full + st, a * 64, head, t * kTileN, b_kv);
    }
  };
  auto load_q = [&]() {
    mbar_expect_tx(bar_q, Cfg::Q_BYTES);
    for (int a = 0; a < Cfg::NATOM; ++a) tma_load_4d(s_q + a * kQAtomBytes, &tm_q, bar_q, a * 64, head, q0, b);
  };

  if (Cfg::SPLIT && warp == 4) {
    // ------------------------------------------------------------ TMA producer (own warp)
    if (FRESCO_ISSUER_THREAD(lane)) {
      load_q();
      for (int t = 0; t < n_tiles; ++t) {
        if (t >= ST) mbar_wait(bar_kv_empty + t % ST, ((t / ST) - 1) & 1, 1);
        load_kv_tile(t);
      }
    }
  } else if (warp == Cfg::QK_WARP) {
    // ------------------------------------------------------------ score-MMA issuer (+ TMA producer when !SPLIT)
    // tcgen05.mma issue is a serial affair for the issuing thread; with 7 small MMAs and 3 commits per 128x64 tile a
    // single issuer thread sat on the critical path.  The score MMAs are therefore issued here and the P V MMAs by
    // another warp.  Without a producer warp of its own this thread also refills the K/V ring, never blocking on it.
    if (FRESCO_ISSUER_THREAD(lane)) {
      constexpr uint32_t idesc_qk = make_idesc_f16(kTileM, kTileN, 0);
      const uint32_t q_addr = smem_u32(s_q);
      int next_load = 0;
      auto refill = [&]() {                     // issue every K/V tile load whose ring stage is free; never blocks
        if (Cfg::SPLIT) return;
        while (next_load < n_tiles) {
          if (next_load >= ST && !mbar_test_wait(bar_kv_empty + next_load % ST, ((next_load / ST) - 1) & 1)) break;
          load_kv_tile(next_load);
          ++next_load;
        }
      };
      auto issue_qk = [&](int t) {
        const int st = t % ST;
        uint32_t polls = 0;
        while (!mbar_try_wait(bar_kv_full + st, (t / ST) & 1)) {   // keep the ring moving while waiting for K_t
          refill();
          if (++polls > FRESCO_WATCHDOG_POLLS) mbar_timeout(bar_kv_full + st, (t / ST) & 1, 10);
        }
        tc_fence_after();
        const uint32_t k_addr = smem_u32(s_kv + st * Cfg::STAGE_BYTES);
        const uint32_t d_tmem = tmem + ((t & 1) ? Cfg::S_OFF1 : Cfg::S_OFF0);
#pragma unroll
        for (int ks = 0; ks < (ABL(p, 32) ? 0 : Cfg::KSTEPS); ++ks) {
          const uint32_t qoff = (ks >> 2) * kQAtomBytes + (ks & 3) * 32;
          const uint32_t koff = (ks >> 2) * kKVAtomBytes + (ks & 3) * 32;
          umma_ss(d_tmem, make_smem_desc_sw128(q_addr + qoff, 16, 1024), make_smem_desc_sw128(k_addr + koff, 16, 1024),
                  idesc_qk, ks > 0);
        }
        umma_commit(bar_s + (t & 1));
        umma_commit(bar_kv_empty + st);          // K_t consumed (second arrival comes from the PV issuer)
      };
      if (!Cfg::SPLIT) load_q();
      refill();
      mbar_wait(bar_q, 0, 11);
      issue_qk(0);
      if (n_tiles > 1) issue_qk(1);
      for (int t = 0; t + 2 < n_tiles; ++t) {
        refill();
        // S buffer (t & 1) is free as soon as the softmax threads hold S_t in registers
        uint32_t polls = 0;
        while (!mbar_try_wait(bar_c + (t & 1), (t >> 1) & 1)) {
          refill();
          if (++polls > FRESCO_WATCHDOG_POLLS) mbar_timeout(bar_c + (t & 1), (t >> 1) & 1, 13);
        }
        TRACE(t, 9, polls);
        issue_qk(t + 2);
        TRACE(t, 10, polls);
      }
      while (!Cfg::SPLIT && next_load < n_tiles) refill();   // (every tile is loaded by now unless n_tiles <= 2)
    }
  } else if (warp == Cfg::PV_WARP) {
    // ------------------------------------------------------------ P V MMA issuer
    if (FRESCO_ISSUER_THREAD(lane)) {
      constexpr uint32_t idesc_pv0 = make_idesc_f16(kTileM, Cfg::N0, 1);
      constexpr uint32_t idesc_pv1 = make_idesc_f16(kTileM, Cfg::N1 > 0 ? Cfg::N1 : 16, 1);
      constexpr uint32_t idesc_ones = make_idesc_f16(kTileM, 16, 1);
      for (int t = 0; t < n_tiles; ++t) {
        const int st = t % ST;
        mbar_wait_backoff(bar_p + (t & 1), (t >> 1) & 1, 20, 12);   // P_t in TMEM
        TRACE(t, 11, st);
        mbar_wait(bar_kv_full + st, (t / ST) & 1, 15);              // V_t landed long ago; observe it for visibility
        tc_fence_after();
        const uint32_t v_addr = smem_u32(s_kv + st * Cfg::STAGE_BYTES + Cfg::NATOM * kKVAtomBytes);
#pragma unroll
        for (int k2 = 0; k2 < (ABL(p, 8) ? 0 : kTileN / 16); ++k2) {
          const uint32_t acc = (k2 > 0 || t > 0) ? 1u : 0u;  // O accumulates in TMEM across all tiles
          const uint32_t p_tmem = tmem + ((t & 1) ? Cfg::P_OFF1 : Cfg::P_OFF0) + k2 * 8;
          umma_ts(tmem + Cfg::O_OFF, p_tmem, make_smem_desc_sw128(v_addr + k2 * 2048, kKVAtomBytes, 1024), idesc_pv0,
                  acc);
          if (Cfg::N1 > 0)
            umma_ts(tmem + Cfg::O_OFF + 64, p_tmem,
                    make_smem_desc_sw128(v_addr + kKVAtomBytes + k2 * 2048, kKVAtomBytes, 1024), idesc_pv1, acc);
          if (Cfg::MMA_ROWSUM)      // l += P_t * ones  (every element of the constant tile is 1.0, so its layout is moot)
            umma_ts(tmem + Cfg::O_OFF + Cfg::L_COL, p_tmem, make_smem_desc_sw128(smem_u32(s_ones), 2048, 1024),
                    idesc_ones, acc);
        }
        umma_commit(bar_kv_empty + st);
        umma_commit(bar_o + (t & 1));
        TRACE(t, 12, st);
      }
    }
  } else if (warp < 4) {
    // ------------------------------------------------------------ softmax warps
    const int row = warp * 32 + lane;                      // query row inside the tile == TMEM lane
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const int q_row = q0 + row;
    const int kv_len = p.kv_len;
    const float scale_log2 = p.scale_log2, bias_log2 = p.diag_bias_log2;
    const bool use_bias = bias_log2 != 0.f;
    const unsigned long long scale2 = pack_f2(scale_log2, scale_log2);
    float m_run = -INFINITY, l_run = 0.f;

    bool s_ready = false;                      // result of the early (overlapped) probe of the next S barrier
    for (int i = 0; i < n_tiles; ++i) {
      const int col0 = i * kTileN;
      // warp-uniform: does this tile need masking (ragged tail) or the diagonal bias?
      const bool special = (col0 + kTileN > kv_len) ||
                           (use_bias && (q0 + warp * 32) < col0 + kTileN && (q0 + warp * 32 + 32) > col0);
      TRACE(i, 0, col0);
      if (!s_ready) mbar_wait(bar_s + (i & 1), (i >> 1) & 1, 2);
      tc_fence_after();
      TRACE(i, 1, col0);
      // ---- the whole 64-column row of scores, once, into registers
      const uint32_t s_addr = t_lane + ((i & 1) ? Cfg::S_OFF1 : Cfg::S_OFF0);
      uint32_t r[64];
      if (!ABL(p, 2)) {
        tmem_ld16(s_addr + 0, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
        tmem_ld16(s_addr + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
        tmem_ld16(s_addr + 32, *reinterpret_cast<uint32_t(*)[16]>(&r[32]));
        tmem_ld16(s_addr + 48, *reinterpret_cast<uint32_t(*)[16]>(&r[48]));
      } else {
#pragma unroll
        for (int j = 0; j < 64; ++j) r[j] = 0x3a83126fu + ((uint32_t)(j ^ i) << 8);   // ~1e-3, no conversions
      }
      tmem_ld_wait_dep64(r);
      TRACE(i, 2, r[63]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_c + (i & 1));   // S buffer (i & 1) may be overwritten by Q K_{i+2}^T
      // P buffer (i & 1) was last read by P_{i-2} V_{i-2}; probe its retirement now, wait (rarely) before the stores
      const bool p_free = (i < 2) || mbar_test_wait(bar_o + (i & 1), ((i - 2) >> 1) & 1);
      // probe S_{i+1} now: it was issued a whole tile ago, and the ~100-cycle latency of a try_wait on an
      // already-completed barrier hides behind the max / exp work instead of opening the next iteration
      s_ready = (i + 1 < n_tiles) && mbar_test_wait(bar_s + ((i + 1) & 1), ((i + 1) >> 1) & 1);
      TRACE(i, 3, (int)s_ready + (int)p_free);
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
      const float m_tile = ABL(p, 16) ? 0.f : fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
      TRACE(i, 4, __float_as_uint(m_tile));
      if (!p_free) {
        mbar_wait(bar_o + (i & 1), ((i - 2) >> 1) & 1, 3);
        tc_fence_after();
      }
      // ---- lazy running max: raise it (and rescale O in TMEM) only when it grows by more than 2^8
      if (i == 0) {
        m_run = m_tile;
      } else {
        const bool need = m_tile > m_run + 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          // O may only be touched once P_{i-1} V_{i-1} has retired (rare path, so the wait is affordable)
          mbar_wait(bar_o + ((i - 1) & 1), ((i - 1) >> 1) & 1, 5);
          tc_fence_after();
          const float alpha = need ? fast_exp2(m_run - m_tile) : 1.0f;
          if (need) {
            l_run *= alpha;
            m_run = m_tile;
          }
#pragma unroll
          for (int c = 0; c < D / 8 + (Cfg::MMA_ROWSUM ? 1 : 0); ++c) {
            uint32_t o[8];
            const bool lchunk = Cfg::MMA_ROWSUM && c == D / 8;                          // last chunk: the row-sum column
            const int col = c * 8;
            const uint32_t addr = t_lane + Cfg::O_OFF +
                                  (lchunk ? Cfg::L_COL : (col < Cfg::N0 ? col : 64 + (col - Cfg::N0)));
            tmem_ld8_sync(addr, o);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
            tmem_st8(addr, o);
          }
        }
      }
      // ---- p = exp2(s*scale - m): all 64 exponentials are issued back to back (nothing volatile in between,
      //      so the MUFU pipe is never left idle waiting for a store), then packed to fp16 into the P region
      const unsigned long long negm2 = pack_f2(-m_run, -m_run);
#pragma unroll
      for (int j = 0; j < 64; j += 2) {
        float t0, t1;
        unpack_f2(fma2(pack_f2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), scale2, negm2), t0, t1);
        if (POLY > 0 && ((j >> 1) % (POLY > 0 ? POLY : 1)) == (POLY - 1)) {
          float p0, p1;
          exp2_poly_x2(t0, t1, p0, p1);                    // FMA-pipe exponential for every POLY-th pair
          r[j] = __float_as_uint(p0);
          r[j + 1] = __float_as_uint(p1);
        } else if (ABL(p, 1)) {
          r[j] = __float_as_uint(t0 * 0.001f);
          r[j + 1] = __float_as_uint(t1 * 0.001f);
        } else {
          r[j] = __float_as_uint(fast_exp2(t0));
          r[j + 1] = __float_as_uint(fast_exp2(t1));
        }
      }
      unsigned long long sum2[4] = {pack_f2(0.f, 0.f), pack_f2(0.f, 0.f), pack_f2(0.f, 0.f), pack_f2(0.f, 0.f)};
      if (!Cfg::MMA_ROWSUM) {
#pragma unroll
        for (int j = 0; j < 64; j += 2)
          sum2[(j >> 1) & 3] = add2(sum2[(j >> 1) & 3], pack_f2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])));
      }
      uint32_t pk[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) pk[j] = pack_half2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
      TRACE(i, 5, pk[31] ^ pk[0] ^ pk[15]);
      const uint32_t p_addr = t_lane + ((i & 1) ? Cfg::P_OFF1 : Cfg::P_OFF0);
      if (!ABL(p, 4)) {
        tmem_st16(p_addr, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
        tmem_st16(p_addr + 16, *reinterpret_cast<uint32_t(*)[16]>(&pk[16]));
      } else if (pk[0] == 0x12345678u && pk[31] == 0x9abcdef0u) {
        tmem_st16(p_addr, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));      // keeps pk live
      }
      TRACE(i, 6, i);
      tmem_st_wait();
      TRACE(i, 7, i);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p + (i & 1));
      TRACE(i, 8, i);
      float sa, sb;
      unpack_f2(add2(add2(sum2[0], sum2[1]), add2(sum2[2], sum2[3])), sa, sb);
      l_run += sa + sb;
    }

    // ---- epilogue: O / l -> fp16 head slice of this row
    mbar_wait(bar_o + ((n_tiles - 1) & 1), ((n_tiles - 1) >> 1) & 1, 4);
    tc_fence_after();
    if (Cfg::MMA_ROWSUM) {
      uint32_t lcol[8];
      tmem_ld8_sync(t_lane + Cfg::O_OFF + Cfg::L_COL, lcol);
      l_run = __uint_as_float(lcol[0]);
    }
    const float inv = 1.f / l_run;
    __half* dst = p.out + (static_cast<size_t>(b) * p.q_len + q_row) * (static_cast<size_t>(p.heads) * D) +
                  static_cast<size_t>(head) * D;
#pragma unroll
    for (int c = 0; c < D / 8; ++c) {
      uint32_t o[8];
      const int col = c * 8;
      tmem_ld8_sync(t_lane + Cfg::O_OFF + (col < Cfg::N0 ? col : 64 + (col - Cfg::N0)), o);
      if (q_row < p.q_len) {
        uint4 pkt;
        pkt.x = pack_half2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
        pkt.y = pack_half2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
        pkt.z = pack_half2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
        pkt.w = pack_half2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
        reinterpret_cast<uint4*>(dst)[c] = pkt;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

// ---------------------------------------------------------------------------------------------
// the "twin" kernel: TWO 128-row query tiles per CTA, 128-key tiles, one CTA per SM, one or two threads per query row
// ---------------------------------------------------------------------------------------------
// Written from the round-2 profiles (profiles/r02_attn_*_hot*.txt): the 64-key kernels execute 368 (pipelined) to 572
// (two threads per row) warp-instructions per 128 x 64 tile of which 161 are the softmax arithmetic itself -- waits,
// arrivals, buffer indexing, the lazy-max test and the polling loops are paid per (warp, tile) whatever the tile holds --
// and the issue slots (56-68 % busy), not the MUFU pipe (65 %), are what every such structure runs into at ~840 clocks
// per tile.  Here the per-tile overhead is paid once per 128 keys, and a softmax warp has only two barrier operations per
// tile (wait for S, announce P): P overwrites the start of the score columns it was computed from (the thread has them
// in registers), and the score MMA that reuses a region is issued behind the P V MMA that reads it by the SAME thread
// (the tensor core keeps the order) -- no P-buffer wait exists, and O is only waited for on the rare lazy-max rescale.
// The two query tiles (A, B) interleave: while the warps of A are in their exponentials the tensor core works for B.
//
// When the O accumulators leave room (head_dim <= 48) the scores rotate through THREE 128-column regions shared by the
// two query tiles (sequence A0 B0 A1 B1 ...: score tile n lives in region n % 3), so Q K^T of a query tile's NEXT key
// tile runs while its warps are still in the exponentials of the current one; with two regions (head_dim 64 / 80) a
// query tile's next scores can only be computed after its P V has read P.
//
// SPLIT = 1: one thread per query row holds the whole 128-score row (setmaxnreg: 224 registers for the softmax
// warpgroups).  SPLIT = 2: two threads per row, 64 keys each -- four softmax warps per sub-partition to cover each
// other's fixed-latency stalls; they agree on ONE running max through a named-barrier OR-reduction and share one
// accumulator, so nothing is merged at the end.
//
//   TMEM  S regions [0,128) [128,256) ([256,384))   fp32 scores; P = fp16 over the start of the same columns
//         O_A, O_B behind them (64 columns each with three regions, 128 with two); row sums: O column head_dim (FOLD)
//         or 16 columns behind O (ones-MMA)
//   The synchronisation protocol is restated and model-checked in tools/twin_protocol_model.py.
template <int D, int SPLIT_>
struct TwinCfg {
  static constexpr int SPLIT = SPLIT_;                    // threads per query row: 1, or 2 (64 keys of a tile each)
  static constexpr int KV = 128;                          // keys per tile
  static constexpr int NATOM = (D + 63) / 64;
  static constexpr int KSTEPS = (D + 15) / 16;
  static constexpr int DPAD = KSTEPS * 16;
  // Row sums l = sum_j p_j come out of the tensor core.  FOLD: V column D (zero-filled by TMA: the tensor map ends at
  // head_dim) is overwritten with 1.0 in shared memory by the two spare warps of the issuer warpgroup before the P V
  // MMAs of a key tile are issued, so O column D IS the row sum and a tile costs no extra MMA.  Needs a spare column
  // inside the one atom that is loaded (head_dim 40); head_dim 64 / 80 multiply P by a constant tile of ones instead.
  static constexpr int PVN = ((D + 1 + 15) / 16) * 16;    // O columns with the folded row sum
#ifdef FRESCO_TWIN_NOFOLD                                   /* A/B measurements only */
  static constexpr bool FOLD = false;
#else
  static constexpr bool FOLD = NATOM == 1 && PVN <= 64;   // (head_dim 80 measured 2.5 % faster with the ones-MMA)
#endif
  static constexpr int OCOLS = FOLD ? PVN : DPAD + 16;
  static constexpr int N0 = (FOLD ? PVN : DPAD) < 64 ? (FOLD ? PVN : DPAD) : 64;
  static constexpr int N1 = (FOLD ? PVN : DPAD) - N0;
  static constexpr int NBUF = (OCOLS <= 64) ? 3 : 2;      // score regions
  static constexpr int O_STRIDE = NBUF == 3 ? 64 : 128;
  static constexpr int O_OFF_A = NBUF * 128, O_OFF_B = O_OFF_A + O_STRIDE;
  static constexpr int L_COL = FOLD ? D : DPAD;           // O column that holds the row sum
  static constexpr int TMEM_COLS = 512;
  static constexpr int TILE_BYTES = NATOM * kQAtomBytes;  // a [128 rows x D] tile: Q tile, K tile or V tile
  static constexpr int STAGES = NATOM == 1 ? 4 : 2;
  static constexpr int STAGE_BYTES = 2 * TILE_BYTES;      // K | V
  static constexpr int SMEM_BYTES = 1024 + 2 * TILE_BYTES + STAGES * STAGE_BYTES + 2048 + 512 + 2048;
  // SPLIT 1: 8 softmax warps + MMA issuer + TMA producer + the two V-patch warps: three whole warpgroups, so that
  // setmaxnreg can move registers from the issuer warpgroup (56 each) to the softmax warpgroups (224 each: a 128-score
  // row plus its packed half live in registers; 2 x 224 + 56 = 3 x 168, the pool the CTA is launched with).
  // SPLIT 2: 16 softmax warps (four per sub-partition, 64 scores each) + the same four: 640 threads x 96 registers.
  static constexpr int SOFTMAX_WARPS = 8 * SPLIT;
  static constexpr int MMA_WARP = SOFTMAX_WARPS, TMA_WARP = SOFTMAX_WARPS + 1, PATCH_WARP0 = SOFTMAX_WARPS + 2;
  static constexpr int THREADS = (SOFTMAX_WARPS + 4) * 32;
  static constexpr int XCH_BYTES = SPLIT == 2 ? 2 * 2 * 128 * 4 : 0;    // tile maxima of the two threads of a row
  // TMEM columns (inside a score region) of the P operand of 16-key step k2: SPLIT 1 packs the 128 keys into columns
  // [0, 64); with SPLIT 2 every thread overwrites the start of its OWN 64 score columns: [0, 32) and [64, 96)
  __host__ __device__ static constexpr int p_col(int k2) { return SPLIT == 1 ? k2 * 8 : (k2 >> 2) * 64 + (k2 & 3) * 8; }
};

__device__ __forceinline__ void tmem_ld_wait_dep128(uint32_t (&r)[128]) {
#define FR8(b) "+r"(r[b]), "+r"(r[b + 1]), "+r"(r[b + 2]), "+r"(r[b + 3]), "+r"(r[b + 4]), "+r"(r[b + 5]), "+r"(r[b + 6]), "+r"(r[b + 7])
  asm volatile("tcgen05.wait::ld.sync.aligned;" : FR8(0), FR8(8), FR8(16), FR8(24) : : "memory");
  asm volatile("" : FR8(32), FR8(40), FR8(48), FR8(56) : : "memory");
  asm volatile("" : FR8(64), FR8(72), FR8(80), FR8(88) : : "memory");
  asm volatile("" : FR8(96), FR8(104), FR8(112), FR8(120) : : "memory");
#undef FR8
}
// Waits of the twin kernel: no out-of-line call on the time-out path.  ptxas caps a whole kernel at its smallest
// setmaxnreg value as soon as the kernel contains a real function call (mbar_timeout's printf), which would leave the
// softmax warps 40 registers instead of 232; the watchdog therefore just traps.
__device__ __forceinline__ void mbar_wait_trap(uint64_t* bar, uint32_t parity) {
  uint32_t polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++polls > FRESCO_WATCHDOG_POLLS) __trap();
  }
}
// one elected lane of a converged warp arrives (no branch, no divergence)
__device__ __forceinline__ void mbar_arrive_elected(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "@P mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}\n" ::"r"(smem_u32(bar))
      : "memory");
}

// OR of `pred` over the 64 threads (two warps) that use named barrier `id`; also a barrier for them
__device__ __forceinline__ bool bar_red_or64(int id, bool pred) {
  uint32_t out;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 q, %2, 0;\n\t"
      "barrier.cta.red.or.pred p, %1, 64, q;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}\n"
      : "=r"(out)
      : "r"(id), "r"((uint32_t)pred)
      : "memory");
  return out != 0;
}
__device__ __forceinline__ void bar_sync64(int id) { asm volatile("barrier.cta.sync %0, 64;" ::"r"(id) : "memory"); }

template <int D, int POLY, int SPLIT>
__global__ void __launch_bounds__(TwinCfg<D, SPLIT>::THREADS, 1)
fresco_attn_twin_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                        const __grid_constant__ CUtensorMap tm_v, const AttnParams p) {
  using Cfg = TwinCfg<D, SPLIT>;
  constexpr int ST = Cfg::STAGES;
  constexpr int KV = Cfg::KV;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_q = smem;                                         // [2][TILE_BYTES]  query tiles A, B
  uint8_t* s_kv = smem + 2 * Cfg::TILE_BYTES;                  // [ST][K tile | V tile]
  uint8_t* s_ones = s_kv + ST * Cfg::STAGE_BYTES;              // [16 keys x 128 B] of fp16 1.0
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ones + 2048);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_kv_full = bars + 1;            // [ST]
  uint64_t* bar_kv_empty = bars + 1 + ST;      // [ST]
  constexpr int NB = Cfg::NBUF;
  uint64_t* bar_s = bars + 1 + 2 * ST;         // [NB] score tile n ready in region n % NB; phase (n / NB) & 1
  uint64_t* bar_p = bar_s + 3;                 // [NB] P of score tile n written by the four warps of its query tile
  uint64_t* bar_pv = bar_s + 6;                // [2]  P V_X(j) retired (O_X stable); phase j & 1 (lazy-max rescale only)
  uint64_t* bar_vp = bar_s + 8;                // [ST] ones column written into V tile t (FOLD)
  // [2] the LAST P V_X retired, completed exactly once.  The epilogue must not use bar_pv for this: with three score
  // regions "S_X(T-1) is ready" only implies that P V_X(T-3) has retired, so bar_pv[X] may be two completions behind
  // when a warp reaches the epilogue, and a parity wait cannot tell "T - 2 completions" from "T" (found with
  // tools/twin_protocol_model.py; on the GPU the margin -- a whole softmax tile -- always hid it)
  uint64_t* bar_fin = bar_s + 8 + ST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 10 + ST);
  float* s_xch = reinterpret_cast<float*>(bar_s + 12 + ST);         // [2 query tiles][2 halves][128 rows] (SPLIT 2)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * kTileM);
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int b_kv = b / p.q_per_kv;
  const int n_tiles = (p.kv_len + KV - 1) / KV;

  if (warp == Cfg::TMA_WARP && lane == 0) {
    mbar_init(bar_q, 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(bar_kv_full + s, 1);
      mbar_init(bar_kv_empty + s, 1);
    }
    for (int k = 0; k < NB; ++k) {
      mbar_init(bar_s + k, 1);
      mbar_init(bar_p + k, 4 * SPLIT);
    }
    mbar_init(bar_pv + 0, 1);
    mbar_init(bar_pv + 1, 1);
    mbar_init(bar_fin + 0, 1);
    mbar_init(bar_fin + 1, 1);
    for (int s = 0; s < ST; ++s) mbar_init(bar_vp + s, 2);
    fence_barrier_init();
  }
  if (warp == Cfg::MMA_WARP) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(&tm_k);
      tma_prefetch_desc(&tm_v);
    }
    __syncwarp();
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  for (int i = threadIdx.x; i < 2048 / 4; i += Cfg::THREADS) reinterpret_cast<uint32_t*>(s_ones)[i] = 0x3C003C00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= Cfg::SOFTMAX_WARPS) {
    // ------------------------------------------------------------ issuer warpgroup
    if (SPLIT == 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == Cfg::TMA_WARP) {
    // ------------------------------------------------------------ TMA producer
    if (FRESCO_ISSUER_THREAD(lane)) {
      mbar_expect_tx(bar_q, 2 * Cfg::TILE_BYTES);
#pragma unroll
      for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int a = 0; a < Cfg::NATOM; ++a)
          tma_load_4d(s_q + x * Cfg::TILE_BYTES + a * kQAtomBytes, &tm_q, bar_q, a * 64, head, q0 + x * kTileM, b);
      for (int t = 0; t < n_tiles; ++t) {
        const int st = t % ST;
        if (t >= ST) mbar_wait_trap(bar_kv_empty + st, ((t / ST) - 1) & 1);
        uint8_t* sk = s_kv + st * Cfg::STAGE_BYTES;
        mbar_expect_tx(bar_kv_full + st, Cfg::STAGE_BYTES);
#pragma unroll
        for (int a = 0; a < Cfg::NATOM; ++a) {
          tma_load_4d(sk + a * kQAtomBytes, &tm_k, bar_kv_full + st, a * 64, head, t * KV, b_kv);
          tma_load_4d(sk + Cfg::TILE_BYTES + a * kQAtomBytes, &tm_v, bar_kv_full + st, a * 64, head, t * KV, b_kv);
        }
      }
    }
  } else if (warp == Cfg::MMA_WARP) {
    // ------------------------------------------------------------ the one MMA issuer.  Score tiles in sequence n = 0, 1, ...
    //      (n = 2 t + x): NB score MMAs up front, then per n: wait for P(n), P(n) V(t) -> O_X, and behind it Q K^T of
    //      score tile n + NB into the region P(n) was read from (same thread: the tensor core keeps the order)
    if (FRESCO_ISSUER_THREAD(lane)) {
      constexpr uint32_t idesc_qk = make_idesc_f16(kTileM, KV, 0);
      constexpr uint32_t idesc_pv0 = make_idesc_f16(kTileM, Cfg::N0, 1);
      constexpr uint32_t idesc_pv1 = make_idesc_f16(kTileM, Cfg::N1 > 0 ? Cfg::N1 : 16, 1);
      constexpr uint32_t idesc_ones = make_idesc_f16(kTileM, 16, 1);
      const uint32_t ones_desc_addr = smem_u32(s_ones);
      // score tile n = 2 * t + x (query tile x, key tile t) lives in region n % NB
      auto issue_pv = [&](int n) {                         // O_X (+)= P(n) V(t), l_X (+)= P(n) 1
        const int x = n & 1, t = n >> 1;
        const uint32_t v_addr = smem_u32(s_kv + (t % ST) * Cfg::STAGE_BYTES + Cfg::TILE_BYTES);
        const uint32_t s_tmem = tmem + (n % NB) * 128;
        const uint32_t o_tmem = tmem + (x ? Cfg::O_OFF_B : Cfg::O_OFF_A);
#pragma unroll
        for (int k2 = 0; k2 < KV / 16; ++k2) {
          const uint32_t acc = (k2 > 0 || t > 0) ? 1u : 0u;
          const uint32_t p_tmem = s_tmem + Cfg::p_col(k2);
          umma_ts(o_tmem, p_tmem, make_smem_desc_sw128(v_addr + k2 * 2048, kQAtomBytes, 1024), idesc_pv0, acc);
          if (Cfg::N1 > 0)
            umma_ts(o_tmem + 64, p_tmem, make_smem_desc_sw128(v_addr + kQAtomBytes + k2 * 2048, kQAtomBytes, 1024),
                    idesc_pv1, acc);
          if (!Cfg::FOLD)
            umma_ts(o_tmem + Cfg::L_COL, p_tmem, make_smem_desc_sw128(ones_desc_addr, 2048, 1024), idesc_ones, acc);
        }
      };
      auto issue_qk = [&](int n) {                         // S(n) = Q_X K(t)^T
        const int x = n & 1, t = n >> 1;
        if (x == 0) {                                      // first use of key tile t
          mbar_wait_trap(bar_kv_full + t % ST, (t / ST) & 1);
          tc_fence_after();
        }
        const uint32_t q_addr = smem_u32(s_q + x * Cfg::TILE_BYTES);
        const uint32_t k_addr = smem_u32(s_kv + (t % ST) * Cfg::STAGE_BYTES);
        const uint32_t d_tmem = tmem + (n % NB) * 128;
#pragma unroll
        for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
          const uint32_t off = (ks >> 2) * kQAtomBytes + (ks & 3) * 32;
          umma_ss(d_tmem, make_smem_desc_sw128(q_addr + off, 16, 1024), make_smem_desc_sw128(k_addr + off, 16, 1024),
                  idesc_qk, ks > 0);
        }
        umma_commit(bar_s + n % NB);
      };
      const int n_total = 2 * n_tiles;
      mbar_wait_trap(bar_q, 0);
      for (int n = 0; n < NB && n < n_total; ++n) issue_qk(n);
      for (int n = 0; n < n_total; ++n) {
        mbar_wait_trap(bar_p + n % NB, (n / NB) & 1);        // P(n) is in TMEM
        if (Cfg::FOLD && !(n & 1)) mbar_wait_trap(bar_vp + (n >> 1) % ST, ((n >> 1) / ST) & 1);   // V(t) has its ones column
        tc_fence_after();
        issue_pv(n);
        umma_commit(bar_pv + (n & 1));
        if (n + 2 >= n_total) umma_commit(bar_fin + (n & 1));   // the last P V of this query tile
        if (n & 1) umma_commit(bar_kv_empty + (n >> 1) % ST);   // K(t), V(t) consumed by both query tiles
        // region n % NB is free once P V(n) has read P: the score MMA issued behind it (same thread, in order) may reuse it
        if (n + NB < n_total) issue_qk(n + NB);
      }
    }
  } else if (Cfg::FOLD) {
    // ------------------------------------------------------------ the two spare warps: the ones column of every V tile
    constexpr int a_ones = D / 64, c_ones = D % 64;              // atom and column inside the atom
    constexpr int chunk = (c_ones * 2) / 16, byte = (c_ones * 2) % 16;
    const int tid = (warp - Cfg::PATCH_WARP0) * 32 + lane;       // 0..63: rows tid and tid + 64
    for (int t = 0; t < n_tiles; ++t) {
      const int st = t % ST;
      mbar_wait_trap(bar_kv_full + st, (t / ST) & 1);            // the TMA writes of V(t) have landed
      uint8_t* v_tile = s_kv + st * Cfg::STAGE_BYTES + Cfg::TILE_BYTES + a_ones * kQAtomBytes;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = tid + rr * 64;                             // 128-byte swizzle: 16-byte chunk index ^= row % 8
        *reinterpret_cast<uint16_t*>(v_tile + r * 128 + ((chunk ^ (r & 7)) << 4) + byte) = 0x3C00u;
      }
      fence_proxy_async_smem();                                  // generic-proxy writes -> visible to the MMA's reads
      __syncwarp();
      mbar_arrive_elected(bar_vp + st);
    }
  }
  } else {
    if constexpr (SPLIT == 1) {
    // ------------------------------------------------------------ softmax warps: query tile X = warp / 4, one row per thread
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int x = warp >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t o_lane = t_lane + (x ? Cfg::O_OFF_B : Cfg::O_OFF_A);
    const int q_row = q0 + x * kTileM + row;
    const int kv_len = p.kv_len;
    const float scale_log2 = p.scale_log2, bias_log2 = p.diag_bias_log2;
    const bool use_bias = bias_log2 != 0.f;
    const unsigned long long scale2 = pack_f2(scale_log2, scale_log2);
    // warp-uniform special tiles: the ragged tail (last tile) and the tile that holds this warp's diagonal
    const int j_tail = (kv_len % KV) ? n_tiles - 1 : n_tiles;
    const int j_diag = use_bias ? (q0 + x * kTileM) / KV : -1;
    float m_run = -INFINITY;
    int buf = x % NB, ph = 0;                              // region and phase of score tile n = 2 j + x

    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t s_lane = t_lane + buf * 128;
      mbar_wait_trap(bar_s + buf, ph);
      tc_fence_after();
      uint32_t r[128];
      tmem_ld32(s_lane + 0, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
      tmem_ld32(s_lane + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
      tmem_ld32(s_lane + 64, *reinterpret_cast<uint32_t(*)[32]>(&r[64]));
      tmem_ld32(s_lane + 96, *reinterpret_cast<uint32_t(*)[32]>(&r[96]));
      tmem_ld_wait_dep128(r);
      if (j >= j_tail || j == j_diag) {                    // rare path: fold mask / bias into the raw scores
        const int col0 = j * KV;
#pragma unroll
        for (int c = 0; c < 128; ++c) {
          float v = __uint_as_float(r[c]);
          if (use_bias && col0 + c == q_row) v += bias_log2 / scale_log2;
          if (col0 + c >= kv_len) v = -INFINITY;
          r[c] = __float_as_uint(v);
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 128; c += 8) {
        mx0 = max3(mx0, __uint_as_float(r[c]), __uint_as_float(r[c + 1]));
        mx1 = max3(mx1, __uint_as_float(r[c + 2]), __uint_as_float(r[c + 3]));
        mx2 = max3(mx2, __uint_as_float(r[c + 4]), __uint_as_float(r[c + 5]));
        mx3 = max3(mx3, __uint_as_float(r[c + 6]), __uint_as_float(r[c + 7]));
      }
      const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
      // ---- lazy running max: raise it (and rescale this row of O and l in TMEM) only when it grows by more than 2^8
      if (j == 0) {
        m_run = m_tile;
      } else {
        const bool need = m_tile > m_run + 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait_trap(bar_pv + x, (j - 1) & 1);           // O_X may only be touched once P V_X(j-1) has retired
          tc_fence_after();
          const float alpha = need ? fast_exp2(m_run - m_tile) : 1.0f;
          if (need) m_run = m_tile;
#pragma unroll
          for (int c = 0; c < Cfg::OCOLS / 8; ++c) {
            uint32_t o[8];
            tmem_ld8_sync(o_lane + c * 8, o);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
            tmem_st8(o_lane + c * 8, o);
          }
        }
      }
      // ---- p = exp2(s * scale - m) -> fp16, 32 keys at a time into columns [0, 64) of the S region
      const unsigned long long negm2 = pack_f2(-m_run, -m_run);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t pk[16];
#pragma unroll
        for (int cc = 0; cc < 32; cc += 2) {
          const int c = g * 32 + cc;
          float t0, t1, e0, e1;
          unpack_f2(fma2(pack_f2(__uint_as_float(r[c]), __uint_as_float(r[c + 1])), scale2, negm2), t0, t1);
          if (POLY > 0 && ((c >> 1) % (POLY > 0 ? POLY : 1)) == (POLY - 1)) {
            exp2_poly_x2(t0, t1, e0, e1);
          } else {
            e0 = fast_exp2(t0);
            e1 = fast_exp2(t1);
          }
          pk[cc >> 1] = pack_half2(e0, e1);
        }
        tmem_st16(s_lane + g * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      mbar_arrive_elected(bar_p + buf);
      // n += 2
      if (NB == 3) {
        if (buf == 0) buf = 2;
        else { buf -= 1; ph ^= 1; }
      } else {
        ph ^= 1;
      }
    }

    // ---- epilogue: O / l -> fp16 head slice of this row
    mbar_wait_trap(bar_fin + x, 0);
    tc_fence_after();
    uint32_t lcol[8];
    tmem_ld8_sync(o_lane + (Cfg::L_COL / 8) * 8, lcol);
    const float inv = 1.f / __uint_as_float(lcol[Cfg::L_COL % 8]);
    __half* dst = p.out + (static_cast<size_t>(b) * p.q_len + q_row) * (static_cast<size_t>(p.heads) * D) +
                  static_cast<size_t>(head) * D;
#pragma unroll
    for (int c = 0; c < D / 8; ++c) {
      uint32_t o[8];
      tmem_ld8_sync(o_lane + c * 8, o);
      if (q_row < p.q_len) {
        uint4 pkt;
        pkt.x = pack_half2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
        pkt.y = pack_half2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
        pkt.z = pack_half2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
        pkt.w = pack_half2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
        reinterpret_cast<uint4*>(dst)[c] = pkt;
      }
    }
    } else {
    // ------------------------------------------------------------ softmax warps, two threads per row: query tile X = warp / 8,
    //      key half = (warp / 4) % 2 (keys [64 half, 64 half + 64) of every tile), rows by warp % 4 (the TMEM lane quarter).
    //      The two threads of a row agree on ONE running max through a named-barrier OR-reduction (one instruction on
    //      the common tile) and share one accumulator; each overwrites the start of its own score columns with its P.
    const int x = warp >> 3, half = (warp >> 2) & 1, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t o_lane = t_lane + (x ? Cfg::O_OFF_B : Cfg::O_OFF_A);
    const int q_row = q0 + x * kTileM + row;
    const int kv_len = p.kv_len;
    const float scale_log2 = p.scale_log2, bias_log2 = p.diag_bias_log2;
    const bool use_bias = bias_log2 != 0.f;
    const unsigned long long scale2 = pack_f2(scale_log2, scale_log2);
    const int pair_bar = 1 + x * 4 + quarter;              // named barrier of the two warps that share these 32 rows
    float* xch_mine = s_xch + (x * 2 + half) * 128 + row;
    float* xch_other = s_xch + (x * 2 + (half ^ 1)) * 128 + row;
    // warp-uniform special tiles of this key half: the ragged tail and the tile that holds this warp's diagonal
    const int tail_num = kv_len - 64 * half - 64;
    const int j_tail = tail_num >= 0 ? tail_num / KV + 1 : 0;
    const int d0 = q0 + x * kTileM + quarter * 32 - 64 * half;
    const int j_diag = (use_bias && d0 >= 0 && (d0 % KV) < 64) ? d0 / KV : -1;
    float m_run = -INFINITY;
    int buf = x % NB, ph = 0;

    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t s_lane = t_lane + buf * 128 + 64 * half;
      mbar_wait_trap(bar_s + buf, ph);
      tc_fence_after();
      uint32_t r[64];
      tmem_ld32(s_lane + 0, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
      tmem_ld32(s_lane + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
      tmem_ld_wait_dep64(r);
      if (j >= j_tail || j == j_diag) {
        const int col0 = j * KV + 64 * half;
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          float v = __uint_as_float(r[c]);
          if (use_bias && col0 + c == q_row) v += bias_log2 / scale_log2;
          if (col0 + c >= kv_len) v = -INFINITY;
          r[c] = __float_as_uint(v);
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 64; c += 8) {
        mx0 = max3(mx0, __uint_as_float(r[c]), __uint_as_float(r[c + 1]));
        mx1 = max3(mx1, __uint_as_float(r[c + 2]), __uint_as_float(r[c + 3]));
        mx2 = max3(mx2, __uint_as_float(r[c + 4]), __uint_as_float(r[c + 5]));
        mx3 = max3(mx3, __uint_as_float(r[c + 6]), __uint_as_float(r[c + 7]));
      }
      const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
      // ---- lazy running max shared by the two threads of a row
      if (bar_red_or64(pair_bar, j == 0 || m_tile > m_run + 8.0f)) {
        *xch_mine = m_tile;
        bar_sync64(pair_bar);
        const float m_new = fmaxf(m_tile, *xch_other);
        const bool need = (j == 0) || (m_new > m_run + 8.0f);
        if (j > 0) {
          mbar_wait_trap(bar_pv + x, (j - 1) & 1);           // O_X may only be touched once P V_X(j-1) has retired
          tc_fence_after();
          const float alpha = need ? fast_exp2(m_run - m_new) : 1.0f;
#pragma unroll
          for (int c = 0; c < Cfg::OCOLS / 8; ++c) {           // the two threads take alternate 8-column chunks of the row
            if ((c & 1) != half) continue;
            uint32_t o[8];
            tmem_ld8_sync(o_lane + c * 8, o);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
            tmem_st8(o_lane + c * 8, o);
          }
        }
        if (need) m_run = m_new;
      }
      // ---- p = exp2(s * scale - m) -> fp16, 32 keys at a time over the start of this thread's own score columns
      const float neg_m = (m_run == -INFINITY) ? 0.f : -m_run;
      const unsigned long long negm2 = pack_f2(neg_m, neg_m);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t pk[16];
#pragma unroll
        for (int cc = 0; cc < 32; cc += 2) {
          const int c = g * 32 + cc;
          float t0, t1, e0, e1;
          unpack_f2(fma2(pack_f2(__uint_as_float(r[c]), __uint_as_float(r[c + 1])), scale2, negm2), t0, t1);
          if (POLY > 0 && ((c >> 1) % (POLY > 0 ? POLY : 1)) == (POLY - 1)) {
            exp2_poly_x2(t0, t1, e0, e1);
          } else {
            e0 = fast_exp2(t0);
            e1 = fast_exp2(t1);
          }
          pk[cc >> 1] = pack_half2(e0, e1);
        }
        tmem_st16(s_lane + g * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      mbar_arrive_elected(bar_p + buf);
      if (NB == 3) {
        if (buf == 0) buf = 2;
        else { buf -= 1; ph ^= 1; }
      } else {
        ph ^= 1;
      }
    }

    // ---- epilogue: O / l -> fp16 head slice; the two threads of a row take alternate 16-byte chunks
    mbar_wait_trap(bar_fin + x, 0);
    tc_fence_after();
    uint32_t lcol[8];
    tmem_ld8_sync(o_lane + (Cfg::L_COL / 8) * 8, lcol);
    const float inv = 1.f / __uint_as_float(lcol[Cfg::L_COL % 8]);
    __half* dst = p.out + (static_cast<size_t>(b) * p.q_len + q_row) * (static_cast<size_t>(p.heads) * D) +
                  static_cast<size_t>(head) * D;
#pragma unroll
    for (int c = 0; c < D / 8; ++c) {
      if ((c & 1) != half) continue;
      uint32_t o[8];
      tmem_ld8_sync(o_lane + c * 8, o);
      if (q_row < p.q_len) {
        uint4 pkt;
        pkt.x = pack_half2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
        pkt.y = pack_half2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
        pkt.z = pack_half2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
        pkt.w = pack_half2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
        reinterpret_cast<uint4*>(dst)[c] = pkt;
      }
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == Cfg::MMA_WARP) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

// {head_dim, heads, tokens, batch} view of a token-major [batch, tokens, heads*head_dim] fp16 tensor.
// cuTensorMapEncodeTiled costs a few microseconds and the same (pointer, shape) comes back every denoise step (torch's
// caching allocator), so encoded maps are kept in a small direct-mapped table.
struct MapKey {
  const void* base;
  int head_dim, heads, tokens, batch, box_rows;
  long long row_stride, batch_stride;
  bool operator==(const MapKey& o) const {
    return base == o.base && head_dim == o.head_dim && heads == o.heads && tokens == o.tokens && batch == o.batch &&
           box_rows == o.box_rows && row_stride == o.row_stride && batch_stride == o.batch_stride;
  }
};
// row_stride / batch_stride in elements (dense: heads*head_dim and tokens*heads*head_dim)
static int make_head_tile_map(CUtensorMap* map, const void* base, int head_dim, int heads, int tokens, int batch,
                              int box_rows, long long row_stride, long long batch_stride) {
  constexpr int kSlots = 64;
  static thread_local MapKey keys[kSlots];
  static thread_local CUtensorMap maps[kSlots];
  static thread_local bool valid[kSlots];
  const MapKey key = {base, head_dim, heads, tokens, batch, box_rows, row_stride, batch_stride};
  const size_t h = (reinterpret_cast<uintptr_t>(base) >> 9) * 0x9E3779B97F4A7C15ull + (size_t)tokens * 31 + box_rows;
  const int slot = (int)((h >> 32) % kSlots);
  if (valid[slot] && keys[slot] == key) {
    *map = maps[slot];
    return 0;
  }
  const cuuint64_t dims[4] = {(cuuint64_t)head_dim, (cuuint64_t)heads, (cuuint64_t)tokens, (cuuint64_t)batch};
  const cuuint64_t strides[3] = {(cuuint64_t)head_dim * 2, (cuuint64_t)row_stride * 2, (cuuint64_t)batch_stride * 2};
  const cuuint32_t box[4] = {64, 1, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const int rc = encode_tiled_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box,
                                  estr, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc == 0) {
    keys[slot] = key;
    maps[slot] = *map;
    valid[slot] = true;
  }
  return rc;
}

// Tuning knobs (fresco_internal.h: option(); environment variable of the same name read once, fresco_set_option()
// overrides).  The defaults are the measured best on B200 per head_dim:
//   FRESCO_ATTN_WIDE    -1 / unset = per head_dim (below); 0 = pipelined kernel (64-key tiles, any head_dim);
//                       1 | 2 = twin kernel with that many threads per query row (head_dim <= 80)
//   FRESCO_ATTN_POLY    every n-th pair of exponentials on the FMA pipe: 0 | 4 | 8; -1 / unset = per kernel (below)
//   FRESCO_ATTN_ROWSUM  pipelined kernel, head_dim 40: row sums from the tensor core
constexpr int kRowsumDefault = 1;
// TFLOP/s measured in isolation at the config-2 shapes (profiles/r02_attn_microbench_*.jsonl; +-2 % run to run):
//   d = 40 (L 4096, Lk 15587): pipelined 445 | twin/1 531, poly8 563, poly4 594 | twin/2 543, poly4 584, poly8 621  [torch SDPA 620]
//   d = 80 (L 1024, Lk 3897):  pipelined 516 | twin/1 628, poly4 679 | twin/2 624, poly4 636, poly8 659              [torch SDPA 854]
//   d = 64 (L 2048, Lk 2048):  pipelined 597 | twin/1 poly4 561 | twin/2 poly4 525                                    [torch SDPA 750]
//   d = 128 (L 1024, Lk 1024): pipelined 510, poly4 543                                                               [torch SDPA 846]
// Built, measured slower and removed again (numbers in DESIGN.md 3.1): a "wide" kernel (2 / 4 threads per row with
// private running max and accumulator per key part: 426 / 343 at d = 40, 587 at d = 80), a "duo" kernel (two threads per
// row sharing one running max, 64-key tiles: 449), a 4-CTA-per-SM kernel without pipelining, prefetched scores, and two
// query tiles taking strict turns on the MUFU pipe.
constexpr int default_rows_split(int head_dim) { return head_dim == 40 ? 2 : (head_dim == 80 ? 1 : 0); }
static int twin_split(int head_dim) {       // 0 = pipelined kernel, 1 | 2 = twin kernel, threads per query row
  int w = option(OPT_ATTN_WIDE, -1);
  if (w < 0) w = default_rows_split(head_dim);
  if (head_dim > 80) return 0;
  return w <= 0 ? 0 : (w >= 2 ? 2 : 1);
}
static int poly_option(int head_dim, int split) {
  return option(OPT_ATTN_POLY, split == 0 ? 0 : ((head_dim == 40 && split == 2) ? 8 : 4));
}

template <int D, int POLY, bool ROWSUM>
static int launch_pipelined(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                            dim3 grid, cudaStream_t stream) {
  using Cfg = AttnCfg<D, ROWSUM>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fresco_attn_kernel<D, POLY, ROWSUM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(attn)");
    attr_set = true;
  }
  fresco_attn_kernel<D, POLY, ROWSUM><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tq, tk, tv, p);
  return check_launch("fresco_attn_kernel");
}

template <int D, int POLY, int SPLIT>
static int launch_twin(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, dim3 grid,
                       cudaStream_t stream) {
  using Cfg = TwinCfg<D, SPLIT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fresco_attn_twin_kernel<D, POLY, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(attn twin)");
    attr_set = true;
  }
  fresco_attn_twin_kernel<D, POLY, SPLIT><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tq, tk, tv, p);
  return check_launch("fresco_attn_twin_kernel");
}

template <int D>
static int launch_attn(const void* q, const void* k, const void* v, void* out, int batch_q, int q_len, int kv_len,
                       int heads, int q_per_kv, long long kv_row_stride, long long kv_batch_stride, float softmax_scale,
                       float diag_bias, cudaStream_t stream) {
  CUtensorMap tq, tk, tv;
  const int batch_kv = batch_q / q_per_kv;
  const long long C = (long long)heads * D;
  const int split = twin_split(D);
  const int kv_box = split > 0 ? 128 : kTileN;                       // key rows per TMA box
  if (make_head_tile_map(&tq, q, D, heads, q_len, batch_q, kTileM, C, C * q_len)) return FRESCO_ERR_TENSORMAP;
  if (make_head_tile_map(&tk, k, D, heads, kv_len, batch_kv, kv_box, kv_row_stride, kv_batch_stride)) return FRESCO_ERR_TENSORMAP;
  if (make_head_tile_map(&tv, v, D, heads, kv_len, batch_kv, kv_box, kv_row_stride, kv_batch_stride)) return FRESCO_ERR_TENSORMAP;
  AttnParams p;
  p.out = static_cast<__half*>(out);
  p.q_len = q_len;
  p.kv_len = kv_len;
  p.heads = heads;
  p.q_per_kv = q_per_kv;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.diag_bias_log2 = diag_bias * 1.4426950408889634f;
  p.ablate = option(OPT_ATTN_ABLATE, 0);
  dim3 grid((q_len + kTileM - 1) / kTileM, heads, batch_q);
  const int poly = poly_option(D, split);
  if constexpr (D <= 80) {
    if (split > 0) {
      dim3 grid2((q_len + 2 * kTileM - 1) / (2 * kTileM), heads, batch_q);
      if (split == 2) {
        if (poly == 4) return launch_twin<D, 4, 2>(tq, tk, tv, p, grid2, stream);
        if (poly == 8) return launch_twin<D, 8, 2>(tq, tk, tv, p, grid2, stream);
        return launch_twin<D, 0, 2>(tq, tk, tv, p, grid2, stream);
      }
      if (poly == 4) return launch_twin<D, 4, 1>(tq, tk, tv, p, grid2, stream);
      if (poly == 8) return launch_twin<D, 8, 1>(tq, tk, tv, p, grid2, stream);
      return launch_twin<D, 0, 1>(tq, tk, tv, p, grid2, stream);
    }
  }
  if constexpr (AttnCfg<D, true>::MMA_ROWSUM) {
    if (option(OPT_ATTN_ROWSUM, kRowsumDefault)) {
      if (poly == 4) return launch_pipelined<D, 4, true>(tq, tk, tv, p, grid, stream);
      if (poly == 8) return launch_pipelined<D, 8, true>(tq, tk, tv, p, grid, stream);
      return launch_pipelined<D, 0, true>(tq, tk, tv, p, grid, stream);
    }
  }
  if (poly == 4) return launch_pipelined<D, 4, false>(tq, tk, tv, p, grid, stream);
  if (poly == 8) return launch_pipelined<D, 8, false>(tq, tk, tv, p, grid, stream);
  return launch_pipelined<D, 0, false>(tq, tk, tv, p, grid, stream);
}

}  // namespace fresco

using namespace fresco;

#ifdef FRESCO_ATTN_TRACE
extern "C" int fresco_debug_attn_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_attn_trace, sizeof(long long) * 32 * 16) == cudaSuccess ? 0 : 1;
}
#endif

// which kernel fresco_attn_fwd launches for a head dim under the current options (bench.py names it in its JSON line)
extern "C" const char* fresco_attn_variant(int head_dim) {
  static thread_local char buf[96];
  const int split = twin_split(head_dim);
  if (split > 0) snprintf(buf, sizeof(buf), "fresco_attn_twin_kernel<%d,poly%d,%d>", head_dim, poly_option(head_dim, split), split);
  else snprintf(buf, sizeof(buf), "fresco_attn_kernel<%d,poly%d> (pipelined)", head_dim, poly_option(head_dim, split));
  return buf;
}

extern "C" int fresco_attn_fwd(const void* q, const void* k, const void* v, void* out, int batch_q, int q_len,
                               int kv_len, int heads, int head_dim, int q_per_kv, float softmax_scale,
                               float diag_bias, void* stream) {
  return fresco_attn_fwd_kv_strided(q, k, v, out, batch_q, q_len, kv_len, heads, head_dim, q_per_kv,
                                    (long long)heads * head_dim, (long long)kv_len * heads * head_dim, softmax_scale,
                                    diag_bias, stream);
}

extern "C" int fresco_attn_fwd_kv_strided(const void* q, const void* k, const void* v, void* out, int batch_q, int q_len,
                                          int kv_len, int heads, int head_dim, int q_per_kv, long long kv_row_stride,
                                          long long kv_batch_stride, float softmax_scale, float diag_bias, void* stream) {
  if (!q || !k || !v || !out) return set_error(FRESCO_ERR_ARG, "fresco_attn_fwd: null pointer");
  if (batch_q <= 0 || q_len <= 0 || kv_len <= 0 || heads <= 0 || q_per_kv <= 0 || batch_q % q_per_kv != 0)
    return set_error(FRESCO_ERR_ARG, "fresco_attn_fwd: bad shape");
  if (kv_row_stride < (long long)heads * head_dim || kv_row_stride % 8 != 0 || kv_batch_stride % 8 != 0 ||
      kv_batch_stride < kv_row_stride)
    return set_error(FRESCO_ERR_ARG, "fresco_attn_fwd: K/V strides must be multiples of 8 elements and cover a row");
  if (softmax_scale <= 0.f) return set_error(FRESCO_ERR_ARG, "fresco_attn_fwd: softmax_scale must be > 0");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define ATTN_CASE(DD)                                                                                                 \
  case DD:                                                                                                             \
    return launch_attn<DD>(q, k, v, out, batch_q, q_len, kv_len, heads, q_per_kv, kv_row_stride, kv_batch_stride,      \
                           softmax_scale, diag_bias, s)
  switch (head_dim) {
    ATTN_CASE(40);
    ATTN_CASE(64);
    ATTN_CASE(80);
    ATTN_CASE(128);
    default: return set_error(FRESCO_ERR_UNSUPPORTED, "fresco_attn_fwd: head_dim must be one of 40, 64, 80, 128");
  }
#undef ATTN_CASE
}
