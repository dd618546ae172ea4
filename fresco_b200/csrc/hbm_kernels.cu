// HBM-bound kernels of the FRESCO hot path: K/V compaction, temporal-guided (FLATTEN)
// attention, bilinear flow warp, the warp_tensor frame chain, the temporal-consistency
// loss forward+backward, Adam, AdaIN and the Gram-loss normalise/transposition.
// All are coalesced / vectorised streaming kernels; none is reshaped into a GEMM.
#include "common.cuh"
#include "fresco_internal.h"

namespace fresco {

// =============================================================================================
// A2  K/V compaction  (src/diffusion_hacked.py:234-247)
// =============================================================================================
__global__ void kv_compact_kernel(const uint4* __restrict__ k, const uint4* __restrict__ v,
                                  const int32_t* __restrict__ idx, uint4* __restrict__ k_out,
                                  uint4* __restrict__ v_out, int chunks, long long rows_per_chunk, int n_sel,
                                  int vec_per_row) {
  const long long total = (long long)chunks * n_sel * vec_per_row;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int part = (int)(t % vec_per_row);
    const long long row = t / vec_per_row;
    const int i = (int)(row % n_sel);
    const int c = (int)(row / n_sel);
    const long long src = ((long long)c * rows_per_chunk + idx[i]) * vec_per_row + part;
    k_out[t] = __ldg(k + src);
    v_out[t] = __ldg(v + src);
  }
}

// =============================================================================================
// A5  temporal-guided attention  (src/diffusion_hacked.py:320-367)
// one CTA per (chunk, trajectory); one warp per head
// =============================================================================================
__global__ void temporal_attn_kernel(const __half* __restrict__ q_raw, const __half* __restrict__ k_raw,
                                     const __half* __restrict__ v_src, __half* __restrict__ out,
                                     const int64_t* __restrict__ fwd_map, const uint8_t* __restrict__ traj_mask,
                                     int frames, int tokens, int heads, int d, float scale) {
  extern __shared__ uint8_t smem_t[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x % tokens;
  const int b = blockIdx.x / tokens;
  const int C = heads * d;
  const int N = frames;
  const int vpr = d / 8;                                       // 16-byte vectors per head row
  // per-warp shared: q,k,v as half [N][d], scores float [N][N]
  const size_t per_warp = (size_t)3 * N * d * sizeof(__half) + (size_t)N * N * sizeof(float);
  uint8_t* base = smem_t + warp * ((per_warp + 15) & ~size_t(15));
  __half* sq = reinterpret_cast<__half*>(base);
  __half* sk = sq + N * d;
  __half* sv = sk + N * d;
  float* sc = reinterpret_cast<float*>(sv + N * d);
  const int h = warp;

  const int nvec = N * vpr;
  for (int t = lane; t < 3 * nvec; t += 32) {
    const int which = t / nvec;
    const int rem = t % nvec;
    const int f = rem / vpr, part = rem % vpr;
    const long long pos = fwd_map[(long long)f * tokens + p];
    const __half* src = which == 0 ? q_raw : (which == 1 ? k_raw : v_src);
    const uint4 val = __ldg(reinterpret_cast<const uint4*>(src + (((long long)b * N + f) * tokens + pos) * C +
                                                           (long long)h * d) + part);
    reinterpret_cast<uint4*>(which == 0 ? sq : (which == 1 ? sk : sv))[rem] = val;
  }
  __syncwarp();
  const uint8_t* mrow = traj_mask + (long long)p * N * N;
  for (int pair = lane; pair < N * N; pair += 32) {
    const int f = pair / N, g = pair % N;
    const uint4* a = reinterpret_cast<const uint4*>(sq + f * d);      // 16-byte shared loads: 8 channels each
    const uint4* bb = reinterpret_cast<const uint4*>(sk + g * d);
    float acc = 0.f;
    for (int c = 0; c < vpr; ++c) {
      const uint4 xa = a[c], yb = bb[c];
      const __half2* xh = reinterpret_cast<const __half2*>(&xa);
      const __half2* yh = reinterpret_cast<const __half2*>(&yb);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 x = __half22float2(xh[u]);
        const float2 y = __half22float2(yh[u]);
        acc = fmaf(x.x, y.x, acc);
        acc = fmaf(x.y, y.y, acc);
      }
    }
    sc[pair] = mrow[pair] ? acc * scale : -INFINITY;
  }
  __syncwarp();
  if (lane < N) {
    float* r = sc + lane * N;
    float m = -INFINITY;
    for (int g = 0; g < N; ++g) m = fmaxf(m, r[g]);
    float s = 0.f;
    for (int g = 0; g < N; ++g) {
      const float e = __expf(r[g] - m);
      r[g] = e;
      s += e;
    }
    const float inv = 1.f / s;
    for (int g = 0; g < N; ++g) r[g] *= inv;
  }
  __syncwarp();
  for (int t = lane; t < nvec; t += 32) {                            // one (frame, 8-channel chunk) per lane
    const int f = t / vpr, part = t % vpr;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int g = 0; g < N; ++g) {
      const float pw = sc[f * N + g];
      const uint4 vv = reinterpret_cast<const uint4*>(sv + g * d)[part];
      const __half2* vh = reinterpret_cast<const __half2*>(&vv);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 y = __half22float2(vh[u]);
        acc[2 * u] = fmaf(pw, y.x, acc[2 * u]);
        acc[2 * u + 1] = fmaf(pw, y.y, acc[2 * u + 1]);
      }
    }
    uint4 o;
    o.x = pack_half2(acc[0], acc[1]);
    o.y = pack_half2(acc[2], acc[3]);
    o.z = pack_half2(acc[4], acc[5]);
    o.w = pack_half2(acc[6], acc[7]);
    reinterpret_cast<uint4*>(sq)[t] = o;                             // q is dead: reuse as the output staging
  }
  __syncwarp();
  for (int t = lane; t < nvec; t += 32) {
    const int f = t / vpr, part = t % vpr;
    const long long pos = fwd_map[(long long)f * tokens + p];
    reinterpret_cast<uint4*>(out + (((long long)b * N + f) * tokens + pos) * C + (long long)h * d)[part] =
        reinterpret_cast<const uint4*>(sq)[t];
  }
}

// =============================================================================================
// W3  bilinear flow warp  (gmflow/geometry.py:41-72)
// =============================================================================================
struct Taps {
  int i00, i01, i10, i11;        // plane offsets (clamped), -1 weight handled through w = 0
  float w00, w01, w10, w11;
};

__device__ __forceinline__ Taps make_taps(float x, float y, int h, int w) {
  Taps t;
  const float xf = floorf(x), yf = floorf(y);
  const float ax = x - xf, ay = y - yf;
  const int x0 = (int)xf, y0 = (int)yf, x1 = x0 + 1, y1 = y0 + 1;
  const bool vx0 = x0 >= 0 && x0 <= w - 1, vx1 = x1 >= 0 && x1 <= w - 1;
  const bool vy0 = y0 >= 0 && y0 <= h - 1, vy1 = y1 >= 0 && y1 <= h - 1;
  const int cx0 = min(max(x0, 0), w - 1), cx1 = min(max(x1, 0), w - 1);
  const int cy0 = min(max(y0, 0), h - 1), cy1 = min(max(y1, 0), h - 1);
  t.i00 = cy0 * w + cx0;
  t.i01 = cy0 * w + cx1;
  t.i10 = cy1 * w + cx0;
  t.i11 = cy1 * w + cx1;
  t.w00 = (vx0 && vy0) ? (1.f - ax) * (1.f - ay) : 0.f;
  t.w01 = (vx1 && vy0) ? ax * (1.f - ay) : 0.f;
  t.w10 = (vx0 && vy1) ? (1.f - ax) * ay : 0.f;
  t.w11 = (vx1 && vy1) ? ax * ay : 0.f;
  // NaN / inf coordinates (never produced by finite flows) sample nothing
  if (!(fabsf(x) < 1e9f) || !(fabsf(y) < 1e9f)) t.w00 = t.w01 = t.w10 = t.w11 = 0.f, t.i00 = t.i01 = t.i10 = t.i11 = 0;
  return t;
}

template <typename PlaneT>
__device__ __forceinline__ float sample_taps(const PlaneT* plane, const Taps& t) {
  return t.w00 * plane[t.i00] + t.w01 * plane[t.i01] + t.w10 * plane[t.i10] + t.w11 * plane[t.i11];
}

__global__ void flow_warp_kernel(const float* __restrict__ src, const float* __restrict__ flow,
                                 float* __restrict__ dst, int batch, int channels, int h, int w, int flow_batch) {
  const int hw = h * w;
  const long long total = (long long)batch * hw;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int pix = (int)(t % hw);
    const int b = (int)(t / hw);
    const int x = pix % w, y = pix / w;
    const float* fl = flow + (long long)(b % flow_batch) * 2 * hw;
    const Taps tp = make_taps(x + fl[pix], y + fl[hw + pix], h, w);
    const float* s = src + (long long)b * channels * hw;
    float* d = dst + (long long)b * channels * hw;
    for (int c = 0; c < channels; ++c) d[(long long)c * hw + pix] = sample_taps(s + (long long)c * hw, tp);
  }
}

// =============================================================================================
// W1  warp_tensor frame chain  (src/flow_utils.py:41-51)
// =============================================================================================
template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_as_float<__half>(const __half* p) { return __half2float(*p); }
template <typename T>
__device__ __forceinline__ void st_from_float(T* p, float v);
template <>
__device__ __forceinline__ void st_from_float<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void st_from_float<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// one CTA per (chunk, channel): the running plane lives in shared memory (fp32, like the
// reference's `latent = sample.to(float32)`), so the N-1 dependent blends are block-local.
template <typename T>
__global__ void warp_chain_smem_kernel(const T* __restrict__ sample, T* __restrict__ out,
                                       const float* __restrict__ bwd_flow, const float* __restrict__ fwd_flow_last,
                                       const float* __restrict__ blend, int frames, int channels, int h, int w) {
  extern __shared__ float planes[];
  const int hw = h * w;
  float* cur = planes;
  float* nxt = planes + hw;
  const int c = blockIdx.x % channels;
  const int j = blockIdx.x / channels;
  auto plane_of = [&](int f) { return ((long long)(j * frames + f) * channels + c) * hw; };

  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    const float v = ld_as_float(sample + plane_of(0) + i);
    cur[i] = v;
    st_from_float(out + plane_of(0) + i, v);
  }
  __syncthreads();
  for (int ii = 0; ii + 1 < frames; ++ii) {
    const float* fl = bwd_flow + (long long)ii * 2 * hw;
    const float* mk = blend + (long long)ii * hw;
    const bool last = (ii + 2 == frames);
#pragma unroll 4
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
      const int x = i % w, y = i / w;
      const Taps tp = make_taps(x + fl[i], y + fl[hw + i], h, w);
      const float m = mk[i];
      const float z = ld_as_float(sample + plane_of(ii + 1) + i);
      const float v = z * (1.f - m) + sample_taps(cur, tp) * m;
      nxt[i] = v;
      if (!last) st_from_float(out + plane_of(ii + 1) + i, v);
    }
    __syncthreads();
    float* t = cur;
    cur = nxt;
    nxt = t;
  }
  // closing blend: frame N-1 <- warp(frame 0, fwd_flow[N-1])  (flow_utils.py:47-51)
  for (int i = threadIdx.x; i < hw; i += blockDim.x) nxt[i] = ld_as_float(sample + plane_of(0) + i);
  __syncthreads();
  const float* mk = blend + (long long)(frames - 1) * hw;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    const int x = i % w, y = i / w;
    const Taps tp = make_taps(x + fwd_flow_last[i], y + fwd_flow_last[hw + i], h, w);
    const float m = mk[i];
    const float v = cur[i] * (1.f - m) + sample_taps(nxt, tp) * m;
    st_from_float(out + plane_of(frames - 1) + i, v);
  }
}

// planes too large for shared memory (image resolution): one launch per chain step, fp32 scratch
__global__ void warp_blend_step_kernel(const float* __restrict__ src_frames, float* __restrict__ dst_frames,
                                       const float* __restrict__ flow, const float* __restrict__ mask, int chunks,
                                       int frames, int channels, int h, int w, int src_f, int dst_f) {
  const int hw = h * w;
  const long long total = (long long)chunks * channels * hw;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t % hw);
    const int c = (int)((t / hw) % channels);
    const int j = (int)(t / ((long long)hw * channels));
    const int x = i % w, y = i / w;
    const Taps tp = make_taps(x + flow[i], y + flow[hw + i], h, w);
    const float m = mask[i];
    const float* s = src_frames + ((long long)(j * frames + src_f) * channels + c) * hw;
    float* d = dst_frames + ((long long)(j * frames + dst_f) * channels + c) * hw;
    d[i] = d[i] * (1.f - m) + sample_taps(s, tp) * m;
  }
}

// =============================================================================================
// O2  temporal-consistency loss forward + backward  (src/diffusion_hacked.py:461-466)
// one CTA per (chunk, channel); frame pairs are walked sequentially with both planes in shared memory.
// The backward of the bilinear warp (the adjoint W^T, a scatter-add in autograd) is evaluated as a GATHER:
// W^T is the same sparse matrix for all chunks*channels planes of a frame pair, so it is built once per batch
// (warp_taps_kernel + a sort on the host) in ELL form -- 8 packed (source:u16, weight:unorm16) slots per
// destination pixel, 32 bytes per row -- and every plane just reads it: no atomics and no data-dependent loop
// in the per-iteration kernel.  Destinations hit by more than 8 taps go to a small overflow list.
// =============================================================================================
__global__ void warp_taps_kernel(const float* __restrict__ flow, int32_t* __restrict__ dest,
                                 float* __restrict__ weight, int frames, int h, int w) {
  const int hw = h * w;
  const long long total = (long long)frames * hw;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t % hw);
    const int f = (int)(t / hw);
    const float* fl = flow + (long long)f * 2 * hw;
    const Taps tp = make_taps((i % w) + fl[i], (i / w) + fl[hw + i], h, w);
    const int idx[4] = {tp.i00, tp.i01, tp.i10, tp.i11};
    const float wt[4] = {tp.w00, tp.w01, tp.w10, tp.w11};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      dest[t * 4 + k] = wt[k] != 0.f ? idx[k] : -1;
      weight[t * 4 + k] = wt[k];
    }
  }
}

__global__ void warp_loss_kernel(const float* __restrict__ cs, const float* __restrict__ fwd_flow,
                                 const float* __restrict__ bwd_flow, const float* __restrict__ fwd_keep,
                                 const float* __restrict__ bwd_keep, const uint4* __restrict__ bwd_ell,
                                 const uint4* __restrict__ fwd_ell, const int32_t* __restrict__ ovf /*[2][frames][n_ovf][3]*/,
                                 int n_ovf, float* __restrict__ grad, float* __restrict__ loss_acc, int accumulate,
                                 int frames, int channels, int h, int w, float k /* 2 / numel */) {
  extern __shared__ float sm[];
  const int hw = h * w;
  float* c1 = sm;
  float* c2 = sm + hw;
  float* s1 = sm + 2 * hw;          // sign(c2 - W_bf c1) * keep_b * k
  float* s2 = sm + 3 * hw;          // sign(c1 - W_ff c2) * keep_f * k
  const int c = blockIdx.x % channels;
  const int b = blockIdx.x / channels;
  auto plane_of = [&](int f) { return ((long long)(b * frames + f) * channels + c) * hw; };
  float loss = 0.f;

  for (int i = threadIdx.x; i < hw; i += blockDim.x) c2[i] = cs[plane_of(0) + i];
  for (int f = 0; f < frames; ++f) {
    const int fn = (f + 1) % frames;
    float* t = c1;
    c1 = c2;                                                    // previous "next" plane becomes c1
    c2 = t;
    __syncthreads();
    for (int i = threadIdx.x; i < hw; i += blockDim.x) c2[i] = cs[plane_of(fn) + i];
    __syncthreads();
    const float* bf = bwd_flow + (long long)f * 2 * hw;
    const float* ff = fwd_flow + (long long)f * 2 * hw;
    const float* mb = bwd_keep + (long long)f * hw;
    const float* mf = fwd_keep + (long long)f * hw;
    // pixel coordinates advance incrementally (no integer division in the loop)
    const int dx = blockDim.x % w, dy = blockDim.x / w;
    int x = threadIdx.x % w, y = threadIdx.x / w;
#pragma unroll 2
    for (int i = threadIdx.x; i < hw; i += blockDim.x, x += dx, y += dy) {
      if (x >= w) {
        x -= w;
        ++y;
      }
      {  // r1 = c2 - W_bf(c1)
        const Taps tp = make_taps(x + bf[i], y + bf[hw + i], h, w);
        const float r = c2[i] - sample_taps(c1, tp);
        const float m = mb[i];
        loss += fabsf(r) * m;
        s1[i] = (r > 0.f ? 1.f : (r < 0.f ? -1.f : 0.f)) * m * k;
      }
      {  // r2 = c1 - W_ff(c2)
        const Taps tp = make_taps(x + ff[i], y + ff[hw + i], h, w);
        const float r = c1[i] - sample_taps(c2, tp);
        const float m = mf[i];
        loss += fabsf(r) * m;
        s2[i] = (r > 0.f ? 1.f : (r < 0.f ? -1.f : 0.f)) * m * k;
      }
    }
    __syncthreads();
    const bool a_add = accumulate || f > 0;                     // frame f   : first touched at f == 0
    const bool b_add = accumulate || f == frames - 1;           // frame f+1 : first touched here, except the wrap to 0
    // adjoint as a gather: 8 packed (source, weight) slots per destination pixel (ELL), two 16-byte loads per row
    const uint4* eb = bwd_ell + (long long)f * hw * 2;
    const uint4* ef = fwd_ell + (long long)f * hw * 2;
    float* ga = grad + plane_of(f);
    float* gb = grad + plane_of(fn);
    auto gather8 = [&](const uint4* ell, int q, const float* sv) {
      const uint4 e0 = __ldg(ell + 2 * q), e1 = __ldg(ell + 2 * q + 1);
      const uint32_t e[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf((float)(e[j] >> 16) * (1.0f / 65535.0f), sv[e[j] & 0xffffu], acc);
      return acc;
    };
    // d/dc1 = s2 - W_bf^T s1      (frame f)
#pragma unroll 4
    for (int q = threadIdx.x; q < hw; q += blockDim.x) {
      const float acc = s2[q] - gather8(eb, q, s1);
      ga[q] = a_add ? ga[q] + acc : acc;
    }
    if (frames == 2) __syncthreads();                           // ga / gb alias the same two planes
    // d/dc2 = s1 - W_ff^T s2      (frame f+1)
#pragma unroll 4
    for (int q = threadIdx.x; q < hw; q += blockDim.x) {
      const float acc = s1[q] - gather8(ef, q, s2);
      gb[q] = b_add ? gb[q] + acc : acc;
    }
    if (n_ovf > 0) {                                            // destinations hit by more than 8 taps (rare)
      __syncthreads();
      const int32_t* ob = ovf + ((long long)0 * frames + f) * n_ovf * 3;
      const int32_t* of = ovf + ((long long)1 * frames + f) * n_ovf * 3;
      for (int e = threadIdx.x; e < n_ovf; e += blockDim.x) {
        if (ob[3 * e] >= 0) atomicAdd(ga + ob[3 * e], -__int_as_float(ob[3 * e + 2]) * s1[ob[3 * e + 1]]);
        if (of[3 * e] >= 0) atomicAdd(gb + of[3 * e], -__int_as_float(of[3 * e + 2]) * s2[of[3 * e + 1]]);
      }
    }
  }
  if (loss_acc != nullptr) {
    __shared__ float red[32];
    loss = warp_sum(loss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = loss;
    __syncthreads();
    if (threadIdx.x < 32) {
      float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
      v = warp_sum(v);
      if (threadIdx.x == 0) atomicAdd(loss_acc, v * k);
    }
  }
}

// =============================================================================================
// O4  Adam  (torch.optim.Adam defaults; src/diffusion_hacked.py:433,485)
// =============================================================================================
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float one_minus_b1, float b2, float one_minus_b2,
                            float step_size, float inv_sqrt_bc2, float eps) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * one_minus_b1;
    const float vi = v[i] * b2 + gi * gi * one_minus_b2;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
  }
}

// =============================================================================================
// O5  AdaIN with the reference's eps quirk  (src/utils.py:58-78)
// one CTA per (sample, channel) plane
// =============================================================================================
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) t = warp_sum(t);
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}

template <typename T>
__global__ void adain_kernel(const float* __restrict__ content, const T* __restrict__ style, T* __restrict__ out,
                             int hw, int content_rounds_to_half) {
  __shared__ float red[32];
  const long long off = (long long)blockIdx.x * hw;
  const float* cp = content + off;
  const T* sp = style + off;
  float cs = 0.f, ss = 0.f;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    float cv = cp[i];
    if (content_rounds_to_half) cv = __half2float(__float2half_rn(cv));   // `cs.data.to(sample.dtype)` (:488)
    cs += cv;
    ss += ld_as_float(sp + i);
  }
  const float c_mean = block_sum(cs, red) / hw;
  const float s_mean = block_sum(ss, red) / hw;
  float cq = 0.f, sq = 0.f;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    float cv = cp[i];
    if (content_rounds_to_half) cv = __half2float(__float2half_rn(cv));
    const float dc = cv - c_mean;
    const float ds = ld_as_float(sp + i) - s_mean;
    cq += dc * dc;
    sq += ds * ds;
  }
  const float c_var = block_sum(cq, red) / (hw - 1);            // unbiased, torch.var default
  const float s_var = block_sum(sq, red) / (hw - 1);
  const float c_std = sqrtf(c_var + 1e-5f);
  const float s_std = sqrtf(s_var + 1.0f);                      // eps slot receives `chunk` = 1 (utils.py:73)
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    float cv = cp[i];
    if (content_rounds_to_half) cv = __half2float(__float2half_rn(cv));
    st_from_float(out + off + i, (cv - c_mean) / c_std * s_std + s_mean);
  }
}

// =============================================================================================
// O3 step 1: row-normalise and transpose to token-major fp16  (src/diffusion_hacked.py:470-473)
// cs [batch, C, L] fp32  ->  xhat [batch, L, C] fp16, norms [batch, L]
// one CTA per (batch, 32 tokens)
// =============================================================================================
__global__ void gram_normalize_kernel(const float* __restrict__ cs, __half* __restrict__ xhat,
                                      float* __restrict__ norms, int channels, int tokens) {
  __shared__ float part[8][33];
  __shared__ float inv_norm[32];
  __shared__ float tile[64][33];
  const int tiles = (tokens + 31) / 32;
  const int l0 = (blockIdx.x % tiles) * 32;
  const int b = blockIdx.x / tiles;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* src = cs + (long long)b * channels * tokens;
  const int l = l0 + lane;
  float ssq = 0.f;
  if (l < tokens)
    for (int c = warp; c < channels; c += 8) {
      const float v = src[(long long)c * tokens + l];
      ssq = fmaf(v, v, ssq);
    }
  part[warp][lane] = ssq;
  __syncthreads();
  if (warp == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += part[i][lane];
    const float nrm = sqrtf(s);
    inv_norm[lane] = 1.f / nrm;
    if (l < tokens) norms[(long long)b * tokens + l] = nrm;
  }
  __syncthreads();
  for (int c0 = 0; c0 < channels; c0 += 64) {
    for (int cc = warp; cc < 64; cc += 8) {
      const int c = c0 + cc;
      tile[cc][lane] = (c < channels && l < tokens) ? src[(long long)c * tokens + l] * inv_norm[lane] : 0.f;
    }
    __syncthreads();
    // write 32 tokens x 64 channels, channel-contiguous
    for (int t = threadIdx.x; t < 32 * 32; t += blockDim.x) {
      const int tok = t / 32, cp = (t % 32) * 2;
      if (l0 + tok < tokens && c0 + cp < channels) {
        const __half2 hv = __floats2half2_rn(tile[cp][tok], tile[cp + 1][tok]);
        *reinterpret_cast<__half2*>(xhat + ((long long)b * tokens + l0 + tok) * channels + c0 + cp) = hv;
      }
    }
    __syncthreads();
  }
}

}  // namespace fresco

using namespace fresco;

static inline int grid_for(long long work, int block, int per_sm = 8) {
  long long g = (work + block - 1) / block;
  const long long cap = (long long)sm_count() * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

extern "C" int fresco_kv_compact(const void* k, const void* v, const int32_t* idx, void* k_out, void* v_out,
                                 int chunks, int rows_per_chunk, int n_sel, int channels, void* stream) {
  if (!k || !v || !idx || !k_out || !v_out) return set_error(FRESCO_ERR_ARG, "fresco_kv_compact: null pointer");
  if (chunks <= 0 || rows_per_chunk <= 0 || n_sel <= 0 || channels <= 0 || channels % 8 != 0)
    return set_error(FRESCO_ERR_ARG, "fresco_kv_compact: bad shape (channels must be a multiple of 8)");
  const int vpr = channels / 8;
  const long long total = (long long)chunks * n_sel * vpr;
  kv_compact_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)k, (const uint4*)v, idx, (uint4*)k_out, (uint4*)v_out, chunks, rows_per_chunk, n_sel, vpr);
  return check_launch("kv_compact_kernel");
}

extern "C" int fresco_temporal_attn_fwd(const void* q_raw, const void* k_raw, const void* v_src, void* out,
                                        const int64_t* fwd_map, const uint8_t* traj_mask, int chunks, int frames,
                                        int tokens, int heads, int head_dim, float scale, void* stream) {
  if (!q_raw || !k_raw || !v_src || !out || !fwd_map || !traj_mask)
    return set_error(FRESCO_ERR_ARG, "fresco_temporal_attn_fwd: null pointer");
  if (chunks <= 0 || frames <= 0 || tokens <= 0 || heads <= 0 || heads > 32 || head_dim % 8 != 0 || frames > 64)
    return set_error(FRESCO_ERR_ARG, "fresco_temporal_attn_fwd: bad shape");
  if (out == v_src) return set_error(FRESCO_ERR_ARG, "fresco_temporal_attn_fwd: out must not alias v_src");
  const size_t per_warp =
      (((size_t)3 * frames * head_dim * sizeof(__half) + (size_t)frames * frames * sizeof(float)) + 15) & ~size_t(15);
  const size_t smem = per_warp * heads;
  if (smem > 200 * 1024) return set_error(FRESCO_ERR_UNSUPPORTED, "fresco_temporal_attn_fwd: frames*head_dim too large");
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(temporal_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(temporal_attn)");
  }
  temporal_attn_kernel<<<chunks * tokens, 32 * heads, smem, (cudaStream_t)stream>>>(
      (const __half*)q_raw, (const __half*)k_raw, (const __half*)v_src, (__half*)out, fwd_map, traj_mask, frames,
      tokens, heads, head_dim, scale);
  return check_launch("temporal_attn_kernel");
}

extern "C" int fresco_flow_warp(const float* src, const float* flow, float* dst, int batch, int channels, int h,
                                int w, int flow_batch, void* stream) {
  if (!src || !flow || !dst) return set_error(FRESCO_ERR_ARG, "fresco_flow_warp: null pointer");
  if (batch <= 0 || channels <= 0 || h <= 0 || w <= 0 || flow_batch <= 0)
    return set_error(FRESCO_ERR_ARG, "fresco_flow_warp: bad shape");
  flow_warp_kernel<<<grid_for((long long)batch * h * w, 256), 256, 0, (cudaStream_t)stream>>>(
      src, flow, dst, batch, channels, h, w, flow_batch);
  return check_launch("flow_warp_kernel");
}

template <typename T>
static int launch_chain(const void* sample, void* out, const float* bwd_flow, const float* fwd_flow_last,
                        const float* blend, int chunks, int frames, int channels, int h, int w, cudaStream_t s) {
  const size_t smem = (size_t)2 * h * w * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(warp_chain_smem_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(warp_chain)");
    attr_set = true;
  }
  warp_chain_smem_kernel<T><<<chunks * channels, 256, smem, s>>>((const T*)sample, (T*)out, bwd_flow, fwd_flow_last,
                                                                 blend, frames, channels, h, w);
  return check_launch("warp_chain_smem_kernel");
}

extern "C" int fresco_warp_fuse_chain(const void* sample, void* out, int is_half, const float* bwd_flow,
                                      const float* fwd_flow_last, const float* blend, int chunks, int frames,
                                      int channels, int h, int w, void* stream) {
  if (!sample || !out || !bwd_flow || !fwd_flow_last || !blend)
    return set_error(FRESCO_ERR_ARG, "fresco_warp_fuse_chain: null pointer");
  if (chunks <= 0 || frames < 2 || channels <= 0 || h <= 0 || w <= 0)
    return set_error(FRESCO_ERR_ARG, "fresco_warp_fuse_chain: bad shape (frames >= 2)");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t plane_bytes = (size_t)h * w * sizeof(float);
  if (2 * plane_bytes <= 200 * 1024) {
    return is_half ? launch_chain<__half>(sample, out, bwd_flow, fwd_flow_last, blend, chunks, frames, channels, h, w, s)
                   : launch_chain<float>(sample, out, bwd_flow, fwd_flow_last, blend, chunks, frames, channels, h, w, s);
  }
  // large planes (image resolution): fp32 only, in place on `out`, one launch per chain step
  if (is_half) return set_error(FRESCO_ERR_UNSUPPORTED, "fresco_warp_fuse_chain: planes > 100 KB need float tensors");
  const long long hw = (long long)h * w;
  const long long total = (long long)chunks * frames * channels * hw;
  if (out != sample) {
    cudaError_t e = cudaMemcpyAsync(out, sample, total * sizeof(float), cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaMemcpyAsync(warp chain)");
  }
  const int grid = grid_for((long long)chunks * channels * hw, 256);
  for (int ii = 0; ii + 1 < frames; ++ii) {
    warp_blend_step_kernel<<<grid, 256, 0, s>>>((const float*)out, (float*)out, bwd_flow + (long long)ii * 2 * hw,
                                                blend + (long long)ii * hw, chunks, frames, channels, h, w, ii, ii + 1);
    int rc = check_launch("warp_blend_step_kernel");
    if (rc) return rc;
  }
  warp_blend_step_kernel<<<grid, 256, 0, s>>>((const float*)out, (float*)out, fwd_flow_last,
                                              blend + (long long)(frames - 1) * hw, chunks, frames, channels, h, w, 0,
                                              frames - 1);
  return check_launch("warp_blend_step_kernel");
}

extern "C" int fresco_warp_taps(const float* flow, int32_t* dest, float* weight, int frames, int h, int w,
                                void* stream) {
  if (!flow || !dest || !weight) return set_error(FRESCO_ERR_ARG, "fresco_warp_taps: null pointer");
  if (frames <= 0 || h <= 0 || w <= 0) return set_error(FRESCO_ERR_ARG, "fresco_warp_taps: bad shape");
  warp_taps_kernel<<<grid_for((long long)frames * h * w, 256), 256, 0, (cudaStream_t)stream>>>(flow, dest, weight,
                                                                                              frames, h, w);
  return check_launch("warp_taps_kernel");
}

extern "C" int fresco_warp_loss_fwd_bwd(const float* cs, const float* fwd_flow, const float* bwd_flow,
                                        const float* fwd_keep, const float* bwd_keep, const void* bwd_ell,
                                        const void* fwd_ell, const int32_t* overflow, int n_overflow, float* grad,
                                        float* loss_acc, int accumulate, int chunks, int frames, int channels, int h,
                                        int w, void* stream) {
  if (!cs || !fwd_flow || !bwd_flow || !fwd_keep || !bwd_keep || !grad || !bwd_ell || !fwd_ell)
    return set_error(FRESCO_ERR_ARG, "fresco_warp_loss_fwd_bwd: null pointer");
  if (n_overflow > 0 && !overflow) return set_error(FRESCO_ERR_ARG, "fresco_warp_loss_fwd_bwd: overflow list missing");
  if (chunks <= 0 || frames < 2 || channels <= 0 || h <= 0 || w <= 0 || h * w > 65535)
    return set_error(FRESCO_ERR_ARG, "fresco_warp_loss_fwd_bwd: bad shape (frames >= 2, h*w <= 65535)");
  const size_t smem = (size_t)4 * h * w * sizeof(float);
  if (smem > 200 * 1024) return set_error(FRESCO_ERR_UNSUPPORTED, "fresco_warp_loss_fwd_bwd: plane too large for shared memory");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(warp_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(warp_loss)");
    attr_set = true;
  }
  const double numel = (double)chunks * frames * channels * h * w;
  warp_loss_kernel<<<chunks * channels, 256, smem, (cudaStream_t)stream>>>(
      cs, fwd_flow, bwd_flow, fwd_keep, bwd_keep, (const uint4*)bwd_ell, (const uint4*)fwd_ell, overflow, n_overflow,
      grad, loss_acc, accumulate, frames, channels, h, w, (float)(2.0 / numel));
  return check_launch("warp_loss_kernel");
}

extern "C" int fresco_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                                int step, float lr, float beta1, float beta2, float eps, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq) return set_error(FRESCO_ERR_ARG, "fresco_adam_step: null pointer");
  if (n <= 0 || step <= 0) return set_error(FRESCO_ERR_ARG, "fresco_adam_step: bad n/step");
  double bc1 = 1.0, bc2 = 1.0, p1 = 1.0, p2 = 1.0;
  for (int i = 0; i < step; ++i) {
    p1 *= (double)beta1;
    p2 *= (double)beta2;
  }
  bc1 = 1.0 - p1;
  bc2 = 1.0 - p2;
  double sq = bc2 > 0 ? 1.0 / __builtin_sqrt(bc2) : 1.0;
  adam_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, 1.f - beta1,
                                                                  beta2, 1.f - beta2, (float)(lr / bc1), (float)sq, eps);
  return check_launch("adam_kernel");
}

extern "C" int fresco_adain(const float* content, const void* style, void* out, int is_half, int planes, int hw,
                            void* stream) {
  if (!content || !style || !out) return set_error(FRESCO_ERR_ARG, "fresco_adain: null pointer");
  if (planes <= 0 || hw <= 1) return set_error(FRESCO_ERR_ARG, "fresco_adain: bad shape");
  if (is_half)
    adain_kernel<__half><<<planes, 256, 0, (cudaStream_t)stream>>>(content, (const __half*)style, (__half*)out, hw, 1);
  else
    adain_kernel<float><<<planes, 256, 0, (cudaStream_t)stream>>>(content, (const float*)style, (float*)out, hw, 0);
  return check_launch("adain_kernel");
}

extern "C" int fresco_gram_normalize(const float* cs, void* xhat, float* norms, int batch, int channels, int tokens,
                                     void* stream) {
  if (!cs || !xhat || !norms) return set_error(FRESCO_ERR_ARG, "fresco_gram_normalize: null pointer");
  if (batch <= 0 || channels <= 0 || tokens <= 0 || channels % 2 != 0)
    return set_error(FRESCO_ERR_ARG, "fresco_gram_normalize: bad shape");
  const int tiles = (tokens + 31) / 32;
  gram_normalize_kernel<<<batch * tiles, 256, 0, (cudaStream_t)stream>>>(cs, (__half*)xhat, norms, channels, tokens);
  return check_launch("gram_normalize_kernel");
}
