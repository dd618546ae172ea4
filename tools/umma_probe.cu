// Bring-up probe for the tcgen05 / TMA building blocks used by the attention kernel.
// Not part of the product: prints layout / numeric diagnostics for each stage so a failure
// in the fused kernel can be localised from one GPU run.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I include \
//          tools/umma_probe.cu fresco_b200/csrc/runtime.cu -o tools/umma_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "../fresco_b200/csrc/common.cuh"
#include "../fresco_b200/csrc/fresco_internal.h"

using namespace fresco;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

// mode 0: dump the TMA tile; mode 1: S = Q K^T (SS); mode 2: O = P V (TS, P from registers)
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k, int mode, int head,
             int row0, int ksteps, int n_pv, uint8_t* dump, float* dout, const __half* p_rows /*[128][128]*/,
             int swap_pack) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_a = smem;
  uint8_t* s_b = smem + 16384;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  uint64_t* bar_mma = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc<256>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 32768);
    tma_load_4d(s_a, &tm_q, bar, 0, head, row0, 0);
    tma_load_4d(s_b, &tm_k, bar, 0, head, row0, 0);
  }
  mbar_wait(bar, 0);
  if (mode == 0) {
    for (int i = threadIdx.x; i < 16384; i += 128) dump[i] = s_a[i];
  } else if (mode == 1) {
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t idesc = make_idesc_f16(128, 128, 0);
      for (int ks = 0; ks < ksteps; ++ks)
        umma_ss(tmem, make_smem_desc_sw128(smem_u32(s_a) + ks * 32, 16, 1024),
                make_smem_desc_sw128(smem_u32(s_b) + ks * 32, 16, 1024), idesc, ks > 0);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, 0);
    tc_fence_after();
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld32(t_lane + c * 32, r);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) dout[threadIdx.x * 128 + c * 32 + i] = __uint_as_float(r[i]);
    }
  } else {
    // P (fp16) -> TMEM columns [0,64), one row per thread
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c = 0; c < 4; ++c) {
      uint32_t pk[16];
      for (int i = 0; i < 16; ++i) {
        const __half a = p_rows[threadIdx.x * 128 + c * 32 + 2 * i];
        const __half b = p_rows[threadIdx.x * 128 + c * 32 + 2 * i + 1];
        __half2 h = swap_pack ? __halves2half2(b, a) : __halves2half2(a, b);
        pk[i] = *reinterpret_cast<uint32_t*>(&h);
      }
      tmem_st16(t_lane + c * 16, pk);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t idesc = make_idesc_f16(128, n_pv, 1);
      for (int k2 = 0; k2 < 8; ++k2)
        umma_ts(tmem + 128, tmem + k2 * 8, make_smem_desc_sw128(smem_u32(s_b) + k2 * 2048, 16384, 1024), idesc,
                k2 > 0);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, 0);
    tc_fence_after();
    for (int c = 0; c < n_pv / 16; ++c) {
      uint32_t r[16];
      tmem_ld16(t_lane + 128 + c * 16, r);
      tmem_ld_wait();
      for (int i = 0; i < 16; ++i) dout[threadIdx.x * 128 + c * 16 + i] = __uint_as_float(r[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

static int make_map(CUtensorMap* map, const void* base, int d, int heads, int tokens, int batch) {
  const cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)heads, (cuuint64_t)tokens, (cuuint64_t)batch};
  const cuuint64_t strides[3] = {(cuuint64_t)d * 2, (cuuint64_t)heads * d * 2, (cuuint64_t)tokens * heads * d * 2};
  const cuuint32_t box[4] = {64, 1, 128, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  return encode_tiled_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                          CU_TENSOR_MAP_SWIZZLE_128B);
}

int main() {
  const int d = 40, heads = 2, tokens = 200, batch = 1, C = heads * d;
  std::vector<__half> hq(tokens * C), hk(tokens * C);
  srand(1);
  for (auto& x : hq) x = __float2half((rand() % 2001 - 1000) / 1000.f);
  for (auto& x : hk) x = __float2half((rand() % 2001 - 1000) / 1000.f);
  __half *dq, *dk, *dp;
  uint8_t* ddump;
  float* dout;
  CK(cudaMalloc(&dq, hq.size() * 2));
  CK(cudaMalloc(&dk, hk.size() * 2));
  CK(cudaMalloc(&ddump, 16384));
  CK(cudaMalloc(&dout, 128 * 128 * 4));
  CK(cudaMalloc(&dp, 128 * 128 * 2));
  CK(cudaMemcpy(dq, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dk, hk.data(), hk.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap tq, tk;
  if (make_map(&tq, dq, d, heads, tokens, batch) || make_map(&tk, dk, d, heads, tokens, batch)) {
    printf("tensor map encode FAILED: %s\n", fresco_last_error());
    return 1;
  }
  printf("tensor maps ok (inner box 64 > head_dim %d)\n", d);
  const int smem = 1024 + 32768 + 64;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int head = 1;

  for (int row0 : {0, 128}) {
    // ---- mode 0: tile layout
    probe_kernel<<<1, 128, smem>>>(tq, tk, 0, head, row0, 3, 48, ddump, dout, dp, 0);
    CK(cudaDeviceSynchronize());
    std::vector<uint8_t> tile(16384);
    CK(cudaMemcpy(tile.data(), ddump, 16384, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int r = 0; r < 128; ++r)
      for (int c = 0; c < 64; ++c) {
        const int chunk = (c / 8) ^ (r % 8);
        const __half got = *reinterpret_cast<__half*>(&tile[r * 128 + chunk * 16 + (c % 8) * 2]);
        const int gr = row0 + r;
        const float exp = (c < d && gr < tokens) ? __half2float(hq[gr * C + head * d + c]) : 0.f;
        if (__half2float(got) != exp) {
          if (bad < 5) printf("  tile mismatch r=%d c=%d got %f exp %f\n", r, c, __half2float(got), exp);
          ++bad;
        }
      }
    printf("[row0=%d] TMA SW128 tile layout + OOB zero fill: %s (%d mismatches)\n", row0, bad ? "FAIL" : "ok", bad);

    // ---- mode 1: S = Q K^T
    probe_kernel<<<1, 128, smem>>>(tq, tk, 1, head, row0, 3, 48, ddump, dout, dp, 0);
    CK(cudaDeviceSynchronize());
    std::vector<float> S(128 * 128);
    CK(cudaMemcpy(S.data(), dout, S.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int i = 0; i < 128; ++i)
      for (int j = 0; j < 128; ++j) {
        double acc = 0;
        if (row0 + i < tokens && row0 + j < tokens)
          for (int c = 0; c < d; ++c)
            acc += (double)__half2float(hq[(row0 + i) * C + head * d + c]) * __half2float(hk[(row0 + j) * C + head * d + c]);
        maxerr = fmax(maxerr, fabs(acc - S[i * 128 + j]));
      }
    printf("[row0=%d] SS MMA  S=QK^T (K-major SW128, 3 k-steps): max err %.3e %s\n", row0, maxerr,
           maxerr < 1e-2 ? "ok" : "FAIL");
    if (maxerr >= 1e-2) printf("   S[0][0..3] = %f %f %f %f\n", S[0], S[1], S[2], S[3]);

    // ---- mode 2: O = P V  (V := the K tensor's tile, MN-major)
    std::vector<__half> hp(128 * 128);
    for (auto& x : hp) x = __float2half((rand() % 1000) / 1000.f);
    CK(cudaMemcpy(dp, hp.data(), hp.size() * 2, cudaMemcpyHostToDevice));
    for (int swap = 0; swap < 2; ++swap) {
      probe_kernel<<<1, 128, smem>>>(tq, tk, 2, head, row0, 3, 48, ddump, dout, dp, swap);
      CK(cudaDeviceSynchronize());
      std::vector<float> Ov(128 * 128);
      CK(cudaMemcpy(Ov.data(), dout, Ov.size() * 4, cudaMemcpyDeviceToHost));
      double me = 0;
      for (int i = 0; i < 128; ++i)
        for (int c = 0; c < d; ++c) {
          double acc = 0;
          for (int j = 0; j < 128; ++j)
            if (row0 + j < tokens) acc += (double)__half2float(hp[i * 128 + j]) * __half2float(hk[(row0 + j) * C + head * d + c]);
          me = fmax(me, fabs(acc - Ov[i * 128 + c]));
        }
      printf("[row0=%d] TS MMA  O=PV (P in TMEM, V MN-major SW128, N=48, pack swap=%d): max err %.3e %s\n", row0, swap,
             me, me < 5e-2 ? "ok" : "FAIL");
    }
  }
  return 0;
}
