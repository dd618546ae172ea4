// Issue-throughput probe of the instructions in the attention softmax loop (clock64, one CTA on one SM).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 tools/pipe_probe.cu -o tools/pipe_probe
// Prints clocks per warp-instruction per SM sub-partition for 1, 2 and 4 warps per sub-partition.
#include <cstdio>
#include <cuda_fp16.h>

#define REP 64
template <int OP>
__global__ void probe(long long* out, float seed) {
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = seed + j * 0.01f + threadIdx.x * 1e-4f; b[j] = seed * 0.5f + j; }
  unsigned acc = 0;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 16; ++it) {
#pragma unroll
    for (int r = 0; r < REP / 8; ++r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[j]));
        if (OP == 1) { unsigned h; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(a[j]), "f"(b[j])); a[j] = __uint_as_float(h); }
        if (OP == 2) { unsigned long long x, y, z; asm volatile("mov.b64 %0, {%3,%4}; mov.b64 %1, {%4,%3}; fma.rn.f32x2 %2, %0, %1, %0; mov.b64 {%3,%4}, %2;" : "=l"(x), "=l"(y), "=l"(z), "+f"(a[j]), "+f"(b[j])); }
        if (OP == 3) { unsigned long long x, y, z; asm volatile("mov.b64 %0, {%3,%4}; mov.b64 %1, {%4,%3}; add.rn.f32x2 %2, %0, %1; mov.b64 {%3,%4}, %2;" : "=l"(x), "=l"(y), "=l"(z), "+f"(a[j]), "+f"(b[j])); }
        if (OP == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(b[j]), "f"(b[(j + 1) & 7]));
        if (OP == 5) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(b[j]), "f"(b[(j + 1) & 7]));
        if (OP == 6) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[j])); unsigned h; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b[j]), "f"(a[j])); b[j] = __uint_as_float(h); }   // does F2FP share the MUFU pipe?
        if (OP == 7) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[j])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(b[j]) : "f"(b[(j + 2) & 7]), "f"(b[(j + 1) & 7])); }
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += a[j] + b[j];
  if (s == 123.456f || acc == 0xdeadbeefu) out[1023] = 1;
  if ((threadIdx.x & 31) == 0) out[threadIdx.x >> 5] = t1 - t0;
}

template <int OP>
void run(const char* name, long long* d) {
  for (int warps_per_sp : {1, 2, 4}) {
    probe<OP><<<1, 128 * warps_per_sp>>>(d, 0.3f);
    cudaDeviceSynchronize();
    probe<OP><<<1, 128 * warps_per_sp>>>(d, 0.3f);
    cudaDeviceSynchronize();
    long long h[16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < 4 * warps_per_sp; ++w) mx = h[w] > mx ? h[w] : mx;
    printf("%-28s warps/SMSP %d: %7.2f clk per warp-instr per SMSP (%lld clk, %d instr/warp)\n", name, warps_per_sp,
           (double)mx / (16.0 * REP * warps_per_sp), mx, 16 * REP);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 1024 * 8);
  run<0>("MUFU.EX2", d);
  run<1>("F2FP.F16.F32.PACK_AB", d);
  run<2>("FFMA2 (+movs)", d);
  run<3>("FADD2 (+movs)", d);
  run<4>("FMNMX3", d);
  run<5>("FFMA", d);
  run<6>("MUFU.EX2 + F2FP pair", d);
  run<7>("MUFU.EX2 + FFMA pair", d);
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
