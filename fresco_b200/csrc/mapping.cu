// M1: pixel correspondence between two frames -- the integer / bit-exact path.
// Replaces get_single_mapping_ind (src/flow_utils.py:57-102): the sequential conflict loop
// (:84-97) keeps, per rounded target, the candidate source with the smallest mse, the earlier
// source winning ties ('>' at :92).  That is the lexicographic arg-min of (mse, source), which a
// 64-bit atomicMin on (float bits << 32 | source) computes in parallel; unlinked targets then
// receive the non-winning sources in ascending order (:99-101) through two prefix sums.
// Arithmetic mirrors the reference's CPU ATen ops bit for bit: the integer-factor bilinear
// downsample is (((a+b)+c)+d)*0.25 over the centre 2x2 block, round is half-to-even, the
// 3-channel mean is ((s0+s1)+s2)/3.
#include "common.cuh"
#include "fresco_internal.h"

namespace fresco {

__device__ __forceinline__ float down4(const float* plane, int Wf, int y, int x, int s) {
  if (s == 1) return plane[(long long)y * Wf + x];
  const int y0 = y * s + s / 2 - 1, x0 = x * s + s / 2 - 1;
  const float a = plane[(long long)y0 * Wf + x0], b = plane[(long long)y0 * Wf + x0 + 1];
  const float c = plane[(long long)(y0 + 1) * Wf + x0], d = plane[(long long)(y0 + 1) * Wf + x0 + 1];
  return __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(a, b), c), d), 0.25f);
}

// block-wide exclusive scan of one int per thread (blockDim.x == 1024)
__device__ int block_excl_scan(int v, int* smem /*[33]*/, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = smem[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    smem[lane] = winc - w;
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  *total = smem[32];
  return smem[warp] + inc - v;
}

__global__ void __launch_bounds__(1024, 1)
mapping_single_kernel(const float* __restrict__ flow /*[2,Hf,Wf] (x,y)*/, const float* __restrict__ occ /*[Hf,Wf]*/,
                      const float* __restrict__ imgs /*[2,3,Hf,Wf]*/, int Hf, int Wf, int s,
                      int64_t* __restrict__ mapping, uint8_t* __restrict__ unlinked,
                      unsigned long long* __restrict__ best, int* __restrict__ used, int* __restrict__ unused_list) {
  __shared__ int sc[33];
  const int H = Hf / s, W = Wf / s, L = H * W;
  const long long plane = (long long)Hf * Wf;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    best[i] = ~0ull;
    used[i] = 0;
  }
  __syncthreads();
  // candidates
  for (int u = threadIdx.x; u < L; u += blockDim.x) {
    const int y = u / W, x = u % W;
    const float fy = __fdiv_rn(down4(flow + plane, Wf, y, x, s), (float)s);     // [[1,0]] swap: (dy, dx)
    const float fx = __fdiv_rn(down4(flow, Wf, y, x, s), (float)s);
    const float wy = rintf(__fadd_rn((float)y, fy));
    const float wx = rintf(__fadd_rn((float)x, fx));
    const bool notocc = !(down4(occ, Wf, y, x, s) > 0.5f);
    const bool ok = wy >= 0.f && wy < (float)H && wx >= 0.f && wx < (float)W && notocc;
    if (ok) {
      const int t = (int)__fadd_rn(__fmul_rn(wy, (float)W), wx);
      const int ty = t / W, tx = t % W;
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = down4(imgs + (3 + c) * plane, Wf, y, x, s);             // source: frame 2
        const float tv = down4(imgs + c * plane, Wf, ty, tx, s);                // target: frame 1
        const float d = __fsub_rn(v, tv);
        acc = __fadd_rn(acc, __fmul_rn(d, d));
      }
      const float mse = __fdiv_rn(acc, 3.0f);
      const unsigned long long key = ((unsigned long long)__float_as_uint(mse) << 32) | (unsigned)u;
      atomicMin(best + t, key);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < L; t += blockDim.x) {
    const unsigned long long b = best[t];
    if (b != ~0ull) {
      const int u = (int)(b & 0xffffffffu);
      mapping[t] = u;
      used[u] = 1;
      unlinked[t] = 0;
    } else {
      unlinked[t] = 1;
    }
  }
  __syncthreads();
  // rank the unused sources / unlinked targets (ascending) and pair them up
  const int per = (L + blockDim.x - 1) / blockDim.x;
  const int lo = min(L, (int)threadIdx.x * per), hi = min(L, lo + per);
  int cnt = 0, total = 0;
  for (int i = lo; i < hi; ++i) cnt += used[i] ? 0 : 1;
  int off = block_excl_scan(cnt, sc, &total);
  for (int i = lo; i < hi; ++i)
    if (!used[i]) unused_list[off++] = i;
  cnt = 0;
  for (int i = lo; i < hi; ++i) cnt += unlinked[i] ? 1 : 0;
  off = block_excl_scan(cnt, sc, &total);        // also orders the unused_list writes (syncthreads inside)
  for (int i = lo; i < hi; ++i)
    if (unlinked[i]) mapping[i] = unused_list[off++];
}

}  // namespace fresco

using namespace fresco;

extern "C" size_t fresco_mapping_workspace_bytes(int tokens) { return (size_t)tokens * 16 + 64; }

extern "C" int fresco_mapping_single(const float* bwd_flow, const float* bwd_occ, const float* imgs, int height,
                                     int width, int scale, int64_t* mapping, uint8_t* unlinked, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  if (!bwd_flow || !bwd_occ || !imgs || !mapping || !unlinked || !workspace)
    return set_error(FRESCO_ERR_ARG, "fresco_mapping_single: null pointer");
  if (scale < 1 || (scale > 1 && scale % 2 != 0) || height % scale != 0 || width % scale != 0)
    return set_error(FRESCO_ERR_UNSUPPORTED, "fresco_mapping_single: scale must be 1 or an even divisor of H and W");
  const int L = (height / scale) * (width / scale);
  if (workspace_bytes < fresco_mapping_workspace_bytes(L))
    return set_error(FRESCO_ERR_ARG, "fresco_mapping_single: workspace too small");
  if (L >= (1 << 24)) return set_error(FRESCO_ERR_UNSUPPORTED, "fresco_mapping_single: too many pixels");
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  unsigned long long* best = reinterpret_cast<unsigned long long*>(ws);
  int* used = reinterpret_cast<int*>(ws + (size_t)8 * L);
  int* unused_list = reinterpret_cast<int*>(ws + (size_t)12 * L);
  mapping_single_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(bwd_flow, bwd_occ, imgs, height, width, scale, mapping,
                                                            unlinked, best, used, unused_list);
  return check_launch("mapping_single_kernel");
}
