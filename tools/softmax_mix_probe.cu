// What does the arithmetic of one softmax tile (32 rows x 64 scores per warp) cost on its own -- no TMEM, no barriers?
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/softmax_mix_probe.cu -o tools/softmax_mix_probe
// Prints SM clocks per tile per sub-partition for 1, 2, 4 warps per sub-partition and several instruction mixes.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>

__device__ __forceinline__ unsigned long long pack_f2(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack_f2(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float max3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) { uint32_t r; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ void exp2_poly_x2(float t0, float t1, float& p0, float& p1) {
  t0 = fmaxf(t0, -126.0f); t1 = fmaxf(t1, -126.0f);
  const unsigned long long t2 = pack_f2(t0, t1), magic = pack_f2(12582912.0f, 12582912.0f);
  const unsigned long long r2 = add2(t2, magic), n2 = add2(r2, pack_f2(-12582912.0f, -12582912.0f));
  const unsigned long long f2 = fma2(n2, pack_f2(-1.0f, -1.0f), t2);
  unsigned long long q2 = fma2(pack_f2(0.0551716f, 0.0551716f), f2, pack_f2(0.2426111f, 0.2426111f));
  q2 = fma2(q2, f2, pack_f2(0.6932610f, 0.6932610f)); q2 = fma2(q2, f2, pack_f2(0.9999281f, 0.9999281f));
  float q0, q1, r0, r1; unpack_f2(q2, q0, q1); unpack_f2(r2, r0, r1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(r0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(r1) << 23));
}

__device__ __forceinline__ void lds4(const uint4* p, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
}
__device__ __forceinline__ void sts4(uint4* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// MIX bits: 1 row max, 2 scale fma2, 4 ex2 (MUFU), 8 row sum add2, 16 fp16 pack, 32 every 4th pair by polynomial
template <int MIX>
__global__ void probe(long long* out, float seed) {
  // scores come from / probabilities go to shared memory (stand-ins for the TMEM load and store), float4 per thread
  // per access with the thread index fastest, so there are no bank conflicts
  extern __shared__ uint4 sm[];
  uint4* s_in = sm;                           // [16][blockDim.x]
  uint4* s_out = sm + 16 * blockDim.x;        // [ 8][blockDim.x]
  for (int j = 0; j < 16; ++j) {
    uint4 v;
    v.x = __float_as_uint(seed * (4 * j + 1) + threadIdx.x * 1e-3f); v.y = __float_as_uint(seed * (4 * j + 2));
    v.z = __float_as_uint(seed * (4 * j + 3)); v.w = __float_as_uint(seed * (4 * j + 4) - threadIdx.x * 1e-3f);
    s_in[j * blockDim.x + threadIdx.x] = v;
  }
  uint32_t r[64];
  float l = 0.f, m_run = 0.f;
  const unsigned long long scale2 = pack_f2(seed, seed);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      lds4(s_in + j * blockDim.x + threadIdx.x, r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    }
    float e[64];
    if (MIX & 1) {
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 64; j += 8) {
        mx0 = max3(mx0, __uint_as_float(r[j]), __uint_as_float(r[j + 1]));
        mx1 = max3(mx1, __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        mx2 = max3(mx2, __uint_as_float(r[j + 4]), __uint_as_float(r[j + 5]));
        mx3 = max3(mx3, __uint_as_float(r[j + 6]), __uint_as_float(r[j + 7]));
      }
      const float mt = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * seed;
      if (mt > m_run + 8.f) m_run = mt;
    }
    const unsigned long long negm2 = pack_f2(-m_run, -m_run);
    unsigned long long sum2[4] = {pack_f2(0.f, 0.f), pack_f2(0.f, 0.f), pack_f2(0.f, 0.f), pack_f2(0.f, 0.f)};
#pragma unroll
    for (int j = 0; j < 64; j += 2) {
      float a0 = __uint_as_float(r[j]), a1 = __uint_as_float(r[j + 1]);
      if (MIX & 2) unpack_f2(fma2(pack_f2(a0, a1), scale2, negm2), a0, a1);
      if ((MIX & 32) && ((j >> 1) & 3) == 3) exp2_poly_x2(a0, a1, a0, a1);
      else if (MIX & 4) { a0 = ex2(a0); a1 = ex2(a1); }
      if (MIX & 8) sum2[(j >> 1) & 3] = add2(sum2[(j >> 1) & 3], pack_f2(a0, a1));
      e[j] = a0; e[j + 1] = a1;
    }
    if (MIX & 16) {
      uint32_t pk[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) pk[j] = pack_half2(e[2 * j], e[2 * j + 1]);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts4(s_out + j * blockDim.x + threadIdx.x, pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)       // unpacked variants: keep every value alive with an fp32 pairwise sum, then store
        sts4(s_out + j * blockDim.x + threadIdx.x, __float_as_uint(e[8 * j] + e[8 * j + 1]),
             __float_as_uint(e[8 * j + 2] + e[8 * j + 3]), __float_as_uint(e[8 * j + 4] + e[8 * j + 5]),
             __float_as_uint(e[8 * j + 6] + e[8 * j + 7]));
    }
    float sa, sb;
    unpack_f2(add2(add2(sum2[0], sum2[1]), add2(sum2[2], sum2[3])), sa, sb);
    l += sa + sb;
  }
  const long long t1 = clock64();
  if (l == 123.456f) out[1023] = 1;
  if ((threadIdx.x & 31) == 0) out[threadIdx.x >> 5] = t1 - t0;
}

template <int MIX>
void run(const char* name, long long* d) {
  cudaFuncSetAttribute(probe<MIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * 16 * 512);
  printf("%-46s", name);
  for (int w : {1, 2, 4}) {
    for (int rep = 0; rep < 2; ++rep) { probe<MIX><<<1, 128 * w, 24 * 16 * 128 * w>>>(d, 0.37f); cudaDeviceSynchronize(); }
    long long h[16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 4 * w; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("  %dw/SMSP: %6.0f clk/tile/warp = %5.0f /tile/SMSP", w, mx / 64.0, mx / 64.0 / w);
  }
  printf("\n");
}

int main() {
  long long* d; cudaMalloc(&d, 1024 * 8);
  run<31>("full: max + fma2 + ex2 + add2 + pack", d);
  run<23>("no row sum (max + fma2 + ex2 + pack)", d);
  run<30>("no row max (fma2 + ex2 + add2 + pack)", d);
  run<6>("fma2 + ex2 only", d);
  run<4>("ex2 only", d);
  run<27>("no ex2 (max + fma2 + add2 + pack)", d);
  run<63>("full, every 4th pair by polynomial", d);
  printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
