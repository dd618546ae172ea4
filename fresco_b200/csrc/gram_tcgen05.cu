// Persistent 128 x 256 tcgen05 GEMM for the spatial-consistency (normalised-Gram L1) loss of optimize_feature
// (src/diffusion_hacked.py:469-476 and its backward), written from the round-2 finding that ISSUING a tcgen05.mma
// costs its thread 60-70 clocks whatever the shape: the 128 x 128 tiles of gemm_tcgen05.cu (N = 128 or 2 x 64 per
// instruction, 32 clocks of tensor work each) were bound by issue, not by the tensor pipe.  Here every instruction is
// M128 N256 K16 (64 clocks of tensor work), one CTA per SM walks a static list of tiles, and the two 256-column TMEM
// accumulators let the four epilogue warps drain tile t while the MMA thread runs tile t + 1.
//
//   MODE_SIGN  D = Xh_I Xh_J^T - Yh_I Yh_J^T   (K = 2C: the second half of the K loop reads the normalised REFERENCE
//              features with the B operand negated by the instruction descriptor), T = 2 sign(D) as fp16 to [I, J] and
//              mirrored to [J, I], loss += sum |D|.  The fp32 [2N, L, L] Gram target of the reference (1.07 GB at layer
//              3, read every Adam iteration) is never materialised: what is stored per batch is Yh, [2N, L, C] fp16.
//              Only tiles that touch the upper triangle are computed (D is symmetric).
//   MODE_GRAD  Ghat = alpha * T Xh   (A = T, K-major; B = Xh read MN-major from its token-major layout), fp32 out.
#include "common.cuh"
#include "fresco_internal.h"

namespace fresco {

constexpr int kG2Stages = 4;
constexpr int kG2AtomA = 128 * 128;                 // [128 rows x 64 fp16]
constexpr int kG2AtomB = 256 * 128;                 // [256 rows x 64 fp16] (K-major) or 4 x [64 k x 64 n] (MN-major)
constexpr int kG2StageBytes = kG2AtomA + kG2AtomB;  // 48 KB
constexpr int kG2Threads = 192;
constexpr int kG2Smem = 1024 + kG2Stages * kG2StageBytes + 256;

enum { G2_SIGN = 0, G2_GRAD = 1 };

struct Gram2Params {
  int M;                      // tokens
  int C;                      // channels
  int batch;
  int T;                      // 128-row tiles per side
  int P;                      // 256-column tile pairs per side
  int tiles_per_batch;
  int n_tiles;                // total
  // G2_SIGN
  __half* tsign;              // [batch, M, M]
  float* loss_acc;
  float loss_scale;           // weight / (batch * M * M)
  // G2_GRAD
  float* out;                 // [batch, M, C]
  float alpha;
};

// tile id -> (batch, row tile I, column pair Jp) for G2_SIGN: row tile I needs the pairs that touch columns >= 128 I,
// i.e. Jp >= I / 2.  G2_GRAD: all pairs of ceil(C / 256).
template <int MODE>
__device__ __forceinline__ void decode_tile(const Gram2Params& p, int tile, int& b, int& I, int& Jp) {
  b = tile / p.tiles_per_batch;
  int r = tile - b * p.tiles_per_batch;
  if (MODE == G2_GRAD) {
    I = r / p.P;
    Jp = r - I * p.P;
    return;
  }
  I = 0;
  for (;;) {
    const int cnt = p.P - (I >> 1);
    if (r < cnt) break;
    r -= cnt;
    ++I;
  }
  Jp = (I >> 1) + r;
}

template <int MODE>
__global__ void __launch_bounds__(kG2Threads, 1)
gram2_kernel(const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_b0,
             const __grid_constant__ CUtensorMap tm_a1, const __grid_constant__ CUtensorMap tm_b1, const Gram2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kG2Stages * kG2StageBytes);
  uint64_t* bar_full = bars;                       // [kG2Stages]
  uint64_t* bar_empty = bars + kG2Stages;          // [kG2Stages]
  uint64_t* bar_acc_full = bars + 2 * kG2Stages;   // [2]
  uint64_t* bar_acc_empty = bar_acc_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // K atoms (64 elements) of one tile: SIGN walks the channels twice (current features, then reference features)
  const int c_atoms = (p.C + 63) / 64;
  const int k_atoms = MODE == G2_SIGN ? 2 * c_atoms : (p.M + 63) / 64;

  if (warp == 5 && lane == 0) {
    for (int s = 0; s < kG2Stages; ++s) {
      mbar_init(bar_full + s, 1);
      mbar_init(bar_empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_acc_full + s, 1);
      mbar_init(bar_acc_empty + s, 4);             // one elected arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_a0);
      tma_prefetch_desc(&tm_b0);
      tma_prefetch_desc(&tm_a1);
      tma_prefetch_desc(&tm_b1);
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (FRESCO_ISSUER_THREAD(lane)) {
      int it = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        int b, I, Jp;
        decode_tile<MODE>(p, tile, b, I, Jp);
        for (int ka = 0; ka < k_atoms; ++ka, ++it) {
          const int st = it % kG2Stages;
          if (it >= kG2Stages) mbar_wait_backoff(bar_empty + st, ((it / kG2Stages) - 1) & 1, 32, 60);
          uint8_t* sa = smem + st * kG2StageBytes;
          uint8_t* sb = sa + kG2AtomA;
          mbar_expect_tx(bar_full + st, kG2StageBytes);
          if (MODE == G2_SIGN) {
            const bool ref = ka >= c_atoms;
            const int kc = (ref ? ka - c_atoms : ka) * 64;
            tma_load_3d(sa, ref ? &tm_a1 : &tm_a0, bar_full + st, kc, I * 128, b);
            tma_load_3d(sb, ref ? &tm_b1 : &tm_b0, bar_full + st, kc, Jp * 256, b);          // [256 rows x 64 k]
          } else {
            tma_load_3d(sa, &tm_a0, bar_full + st, ka * 64, I * 128, b);                       // T rows, K = tokens
#pragma unroll
            for (int q = 0; q < 4; ++q)                                                        // Xh [64 k x 64 n] x 4
              tma_load_3d(sb + q * 8192, &tm_b0, bar_full + st, Jp * 256 + q * 64, ka * 64, b);
          }
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    if (FRESCO_ISSUER_THREAD(lane)) {
      constexpr uint32_t idesc = make_idesc_f16(128, 256, MODE == G2_GRAD ? 1 : 0);
      constexpr uint32_t idesc_negb = idesc | (1u << 14);                                      // B operand negated
      int it = 0, t = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++t) {
        const int buf = t & 1;
        if (t >= 2) mbar_wait(bar_acc_empty + buf, ((t >> 1) - 1) & 1, 61);
        tc_fence_after();
        const uint32_t d_tmem = tmem + buf * 256;
        for (int ka = 0; ka < k_atoms; ++ka, ++it) {
          const int st = it % kG2Stages;
          mbar_wait(bar_full + st, (it / kG2Stages) & 1, 62);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + st * kG2StageBytes);
          const uint32_t b_addr = a_addr + kG2AtomA;
          const uint32_t id = (MODE == G2_SIGN && ka >= c_atoms) ? idesc_negb : idesc;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t acc = (ka > 0 || kk > 0) ? 1u : 0u;
            const uint64_t a_desc = make_smem_desc_sw128(a_addr + kk * 32, 16, 1024);
            const uint64_t b_desc = MODE == G2_GRAD ? make_smem_desc_sw128(b_addr + kk * 2048, 8192, 1024)
                                                    : make_smem_desc_sw128(b_addr + kk * 32, 16, 1024);
            umma_ss(d_tmem, a_desc, b_desc, id, acc);
          }
          umma_commit(bar_empty + st);
        }
        umma_commit(bar_acc_full + buf);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (row = TMEM lane)
    const int row = threadIdx.x;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    float loss = 0.f;
    int t = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++t) {
      int b, I, Jp;
      decode_tile<MODE>(p, tile, b, I, Jp);
      const int buf = t & 1;
      const int gm = I * 128 + row;
      mbar_wait(bar_acc_full + buf, (t >> 1) & 1, 63);
      tc_fence_after();
      const uint32_t acc = t_lane + buf * 256;
      if (MODE == G2_GRAD) {
        float* dst = p.out + ((size_t)b * p.M + gm) * p.C + Jp * 256;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          uint32_t r[32];
          tmem_ld32(acc + c * 32, r);
          tmem_ld_wait();
          if (gm < p.M) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              if (Jp * 256 + c * 32 + i < p.C) {
                float4 v = make_float4(__uint_as_float(r[i]) * p.alpha, __uint_as_float(r[i + 1]) * p.alpha,
                                       __uint_as_float(r[i + 2]) * p.alpha, __uint_as_float(r[i + 3]) * p.alpha);
                *reinterpret_cast<float4*>(dst + c * 32 + i) = v;
              }
            }
          }
        }
      } else {
        const size_t plane = (size_t)b * p.M * p.M;
        __half* dst = p.tsign + plane + (size_t)gm * p.M;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {                      // the two 128-column blocks J of the pair
          const int J = 2 * Jp + half;
          if (J < I || J >= p.T) continue;                          // below the diagonal: the mirror of tile (J, I) covers it
          const bool diag = (J == I);
          const float wgt = diag ? 1.f : 2.f;                       // off-diagonal blocks stand for their mirror as well
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            tmem_ld32(acc + half * 128 + c * 32, r);
            tmem_ld_wait();
            const int j0 = J * 128 + c * 32;
            uint32_t pk[16];
            float part = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float d0 = __uint_as_float(r[i]), d1 = __uint_as_float(r[i + 1]);
              part += fabsf(d0) + fabsf(d1);
              const float t0 = d0 > 0.f ? 2.f : (d0 < 0.f ? -2.f : 0.f);
              const float t1 = d1 > 0.f ? 2.f : (d1 < 0.f ? -2.f : 0.f);
              pk[i >> 1] = pack_half2(t0, t1);
            }
            // columns past M are zero (TMA zero fill) and their T is never read; rows past M are skipped
            if (gm < p.M) {
              // ragged tail: only whole 8-column groups inside M are stored (M % 8 == 0 is checked on the host)
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4)
                if (j0 + q4 * 8 < p.M)
                  *reinterpret_cast<uint4*>(dst + j0 + q4 * 8) =
                      make_uint4(pk[q4 * 4], pk[q4 * 4 + 1], pk[q4 * 4 + 2], pk[q4 * 4 + 3]);
              // columns >= M contribute |0| = 0
              loss += part * wgt;
              if (!diag) {
                // mirror T[j, i] = T[i, j]: for a fixed column j the 32 lanes of a warp write 64 contiguous bytes
                __half* mdst = p.tsign + plane + gm;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  const int j = j0 + i;
                  if (j < p.M) {
                    const uint32_t w = pk[i >> 1];
                    const unsigned short hv = (i & 1) ? (unsigned short)(w >> 16) : (unsigned short)(w & 0xffffu);
                    *reinterpret_cast<unsigned short*>(mdst + (size_t)j * p.M) = hv;
                  }
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + buf);
    }
    if (MODE == G2_SIGN && p.loss_acc != nullptr) {
      loss = warp_sum(loss);
      if (lane == 0) atomicAdd(p.loss_acc, loss * p.loss_scale);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<512>(tmem);
}

// K-major operand [batch, rows, K] fp16 -> {K, rows, batch}, box {64, box_rows, 1}
static int make_kmajor_map2(CUtensorMap* map, const void* base, int rows, int K, int batch, int box_rows) {
  const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
  const cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)rows * K * 2};
  const cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  return encode_tiled_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                          CU_TENSOR_MAP_SWIZZLE_128B);
}
// MN-major B operand [batch, K rows, N cols] fp16 -> {N, K, batch}, box {64, 64, 1}
static int make_mnmajor_map2(CUtensorMap* map, const void* base, int K, int N, int batch) {
  const cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)K, (cuuint64_t)batch};
  const cuuint64_t strides[2] = {(cuuint64_t)N * 2, (cuuint64_t)K * N * 2};
  const cuuint32_t box[3] = {64, 64, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  return encode_tiled_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                          CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int MODE>
static int launch_gram2(const CUtensorMap& a0, const CUtensorMap& b0, const CUtensorMap& a1, const CUtensorMap& b1,
                        const Gram2Params& p, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gram2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kG2Smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(gram2)");
    attr_set = true;
  }
  int grid = sm_count();
  if (grid > p.n_tiles) grid = p.n_tiles;
  gram2_kernel<MODE><<<grid, kG2Threads, kG2Smem, s>>>(a0, b0, a1, b1, p);
  return check_launch("gram2_kernel");
}

}  // namespace fresco

using namespace fresco;

extern "C" int fresco_gram_sign_ref(const void* xhat, const void* yhat, void* tsign, float* loss_acc, int batch,
                                    int tokens, int channels, float weight, void* stream) {
  if (!xhat || !yhat || !tsign) return set_error(FRESCO_ERR_ARG, "fresco_gram_sign_ref: null pointer");
  if (batch <= 0 || tokens <= 0 || channels <= 0 || channels % 8 != 0 || tokens % 8 != 0)
    return set_error(FRESCO_ERR_ARG, "fresco_gram_sign_ref: tokens and channels must be multiples of 8");
  CUtensorMap ax, bx, ay, by;
  if (make_kmajor_map2(&ax, xhat, tokens, channels, batch, 128)) return FRESCO_ERR_TENSORMAP;
  if (make_kmajor_map2(&bx, xhat, tokens, channels, batch, 256)) return FRESCO_ERR_TENSORMAP;
  if (make_kmajor_map2(&ay, yhat, tokens, channels, batch, 128)) return FRESCO_ERR_TENSORMAP;
  if (make_kmajor_map2(&by, yhat, tokens, channels, batch, 256)) return FRESCO_ERR_TENSORMAP;
  Gram2Params p = {};
  p.M = tokens;
  p.C = channels;
  p.batch = batch;
  p.T = (tokens + 127) / 128;
  p.P = (p.T + 1) / 2;
  int per = 0;
  for (int I = 0; I < p.T; ++I) per += p.P - (I >> 1);
  p.tiles_per_batch = per;
  p.n_tiles = per * batch;
  p.tsign = static_cast<__half*>(tsign);
  p.loss_acc = loss_acc;
  p.loss_scale = (float)((double)weight / ((double)batch * tokens * tokens));
  return launch_gram2<G2_SIGN>(ax, bx, ay, by, p, (cudaStream_t)stream);
}

extern "C" int fresco_gram_tx(const void* tsign, const void* xhat, float* ghat, int batch, int tokens, int channels,
                              float alpha, void* stream) {
  if (!tsign || !xhat || !ghat) return set_error(FRESCO_ERR_ARG, "fresco_gram_tx: null pointer");
  if (batch <= 0 || tokens <= 0 || channels <= 0 || channels % 8 != 0 || tokens % 8 != 0)
    return set_error(FRESCO_ERR_ARG, "fresco_gram_tx: tokens and channels must be multiples of 8");
  CUtensorMap ta, tb;
  if (make_kmajor_map2(&ta, tsign, tokens, tokens, batch, 128)) return FRESCO_ERR_TENSORMAP;
  if (make_mnmajor_map2(&tb, xhat, tokens, channels, batch)) return FRESCO_ERR_TENSORMAP;
  Gram2Params p = {};
  p.M = tokens;
  p.C = channels;
  p.batch = batch;
  p.T = (tokens + 127) / 128;
  p.P = (channels + 255) / 256;
  p.tiles_per_batch = p.T * p.P;
  p.n_tiles = p.tiles_per_batch * batch;
  p.out = ghat;
  p.alpha = alpha;
  return launch_gram2<G2_GRAD>(ta, tb, ta, tb, p, (cudaStream_t)stream);
}
