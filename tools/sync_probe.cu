// Latency probe of the synchronisation primitives used by the attention pipeline (clock64, one CTA).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I include tools/sync_probe.cu -o tools/sync_probe
#include <cstdio>
#include "../fresco_b200/csrc/common.cuh"
using namespace fresco;

__global__ void probe(long long* out) {
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t slot;
  __shared__ volatile int flag;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar + 0, 1);
    mbar_init(bar + 1, 1);
    mbar_init(bar + 2, 32);
    mbar_init(bar + 3, 1);
    fence_barrier_init();
    flag = 0;
  }
  if (warp == 0) tmem_alloc<64>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int R = 64;
  long long t0, t1;
  if (threadIdx.x == 0) {
    // (a) tcgen05.commit with nothing pending -> own spin on test_wait
    t0 = clock64();
    for (int i = 0; i < R; ++i) {
      umma_commit(bar + 0);
      while (!mbar_test_wait(bar + 0, i & 1)) {}
    }
    t1 = clock64();
    out[0] = (t1 - t0) / R;
    // (b) plain mbarrier.arrive -> own spin on test_wait
    t0 = clock64();
    for (int i = 0; i < R; ++i) {
      mbar_arrive(bar + 1);
      while (!mbar_test_wait(bar + 1, i & 1)) {}
    }
    t1 = clock64();
    out[1] = (t1 - t0) / R;
    // (c) fences and tcgen05 waits with nothing pending
    t0 = clock64();
    for (int i = 0; i < R; ++i) { tc_fence_before(); tc_fence_after(); }
    t1 = clock64();
    out[2] = (t1 - t0) / R;
  }
  __syncthreads();
  if (warp == 0) {
    t0 = clock64();
    for (int i = 0; i < R; ++i) tmem_ld_wait();
    t1 = clock64();
    if (lane == 0) out[3] = (t1 - t0) / R;
    t0 = clock64();
    for (int i = 0; i < R; ++i) tmem_st_wait();
    t1 = clock64();
    if (lane == 0) out[4] = (t1 - t0) / R;
    // (d) tcgen05.ld x16 + wait
    uint32_t r[16];
    t0 = clock64();
    for (int i = 0; i < R; ++i) { tmem_ld16(slot, r); tmem_ld_wait(); }
    t1 = clock64();
    if (lane == 0) out[5] = (t1 - t0) / R + (r[0] == 12345 ? 1 : 0);
    // (e) tcgen05.st x16 + wait
    t0 = clock64();
    for (int i = 0; i < R; ++i) { tmem_st16(slot, r); tmem_st_wait(); }
    t1 = clock64();
    if (lane == 0) out[6] = (t1 - t0) / R;
  }
  __syncthreads();
  // (f) ping-pong between two warps through mbarriers: warp 1 arrives (32 lanes) on bar2, warp 2 lane 0 arrives on bar3
  if (warp == 1) {
    t0 = clock64();
    for (int i = 0; i < R; ++i) {
      mbar_arrive(bar + 2);
      while (!mbar_test_wait(bar + 3, i & 1)) {}
    }
    t1 = clock64();
    if (lane == 0) out[7] = (t1 - t0) / R;
  } else if (warp == 2 && lane == 0) {
    for (int i = 0; i < R; ++i) {
      while (!mbar_test_wait(bar + 2, i & 1)) {}
      mbar_arrive(bar + 3);
    }
  }
  __syncthreads();
  // (g) same ping-pong but the reply is a tcgen05.commit (as the MMA thread does) and waits use try_wait
  if (threadIdx.x == 0) { mbar_init(bar + 2, 32); mbar_init(bar + 3, 1); fence_barrier_init(); }
  __syncthreads();
  if (warp == 1) {
    t0 = clock64();
    for (int i = 0; i < R; ++i) {
      tc_fence_before();
      mbar_arrive(bar + 2);
      while (!mbar_try_wait(bar + 3, i & 1)) {}
      tc_fence_after();
    }
    t1 = clock64();
    if (lane == 0) out[8] = (t1 - t0) / R;
  } else if (warp == 2 && lane == 0) {
    for (int i = 0; i < R; ++i) {
      while (!mbar_try_wait(bar + 2, i & 1)) {}
      tc_fence_after();
      umma_commit(bar + 3);
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(slot);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16 * sizeof(long long));
  cudaMemset(d, 0, 16 * sizeof(long long));
  probe<<<1, 96>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[] = {"tcgen05.commit (idle) -> own test_wait", "mbarrier.arrive -> own test_wait",
                         "fence before+after", "tcgen05.wait::ld (idle)", "tcgen05.wait::st (idle)",
                         "tcgen05.ld x16 + wait", "tcgen05.st x16 + wait", "warp<->thread ping-pong (arrive/test_wait)",
                         "ping-pong with commit reply + try_wait + fences"};
  printf("status %s\n", cudaGetErrorString(e));
  for (int i = 0; i < 9; ++i) printf("%-52s %6lld cycles\n", names[i], h[i]);
  return 0;
}
