// kernels_template.cuh -- key-frame work: TemplateData::setData on the device (sm_100a).
//
// Replaces (reference file:line):
//   saliency map incl. its live indexing bugs    bpvo/dense_descriptor.cc:92-100, bpvo/imgproc.cc:45-127
//   IsLocalMax (3x4 SIMD window for r = 1)        bpvo/imgproc.h:93-165
//   scan-order selection + disparity gate         bpvo/template_data.cc:51-83
//   makePoint                                     bpvo/rigid_body_warp.h:47-60
//   multiple-of-16 trim                           bpvo/template_data.cc:85-89
//   Hartley normalisation                         bpvo/warps.cc:27-48
//   I0 gather + central-difference gradients      bpvo/template_data.cc:105-131
//
// Selection order is the reference's scan order (y, then x) -- an ORDERED stream compaction:
// flags + per-block counts -> single-CTA exclusive scan -> scatter.
#pragma once

#include "device_types.h"

namespace bp {

// descriptor value of channel c at linear pixel p (interleaved layout)
template <int C> __device__ __forceinline__ float dval(const float* __restrict__ d, int p, int c) {
  return __ldg(d + (size_t) p * kStride<C> + c);
}

// gradientAbsMag at linear pixel p of channel c: abs(I[p-1]-I[p+1]) + abs(I[p-cols]-I[p+cols]).
// Linear indexing reproduces the reference's row wrap-around at columns 0 and cols-1.
template <int C> __device__ __forceinline__ float grad_mag(const float* __restrict__ d, int p, int cols, int c) {
  const float ix = fabsf(__fsub_rn(dval<C>(d, p - 1, c), dval<C>(d, p + 1, c)));
  const float iy = fabsf(__fsub_rn(dval<C>(d, p - cols, c), dval<C>(d, p + cols, c)));
  return __fadd_rn(ix, iy);
}
// the scalar tails of imgproc.cc:64-66 / :120-122 ADD the vertical neighbours (sic, Q4)
template <int C> __device__ __forceinline__ float grad_mag_tail(const float* __restrict__ d, int p, int cols, int c) {
  const float ix = fabsf(__fsub_rn(dval<C>(d, p + 1, c), dval<C>(d, p - 1, c)));
  const float iy = fabsf(__fadd_rn(dval<C>(d, p + cols, c), dval<C>(d, p - cols, c)));
  return __fadd_rn(ix, iy);
}

// Saliency map in "reference compatible" form.  Closed form of what the reference's two functions
// leave in memory:
//   rows 0 and R-1: 0;  column cols-1: 0
//   C == 1 : col < n: gradmag(ch0), col >= n: tail form        (n = cols & ~3)
//   C  > 1 : 4 <= col < n: gradmag(ch0) ONLY (store-to-dst bug), col >= n: sum of tails over all channels,
//            col < 4: S0[n-4+col] + gradmag(ch C-1)[n-4+col]  (the last 4-wide store of the last channel)
template <int C>
__global__ void __launch_bounds__(256) saliency_kernel(const float* __restrict__ d, int rows, int cols, float* __restrict__ S) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= rows * cols) return;
  const int y = p / cols, x = p - y * cols;
  float v = 0.0f;
  if (y > 0 && y < rows - 1 && x != cols - 1) {
    const int n = cols & ~3;
    if (C == 1) {
      v = (x < n) ? grad_mag<C>(d, p, cols, 0) : grad_mag_tail<C>(d, p, cols, 0);
    } else if (x >= n) {
      v = grad_mag_tail<C>(d, p, cols, 0);
      for (int c = 1; c < C; ++c) v = __fadd_rn(v, grad_mag_tail<C>(d, p, cols, c));
    } else if (x >= 4) {
      v = grad_mag<C>(d, p, cols, 0);
    } else {
      const int j = n - 4 + x, pj = y * cols + j;
      const float base = (j == cols - 1) ? 0.0f : grad_mag<C>(d, pj, cols, 0);
      v = __fadd_rn(base, grad_mag<C>(d, pj, cols, C - 1));
    }
  }
  S[p] = v;
}

struct SelectArgs {
  const float* S;        // saliency [rows][cols]
  const float* D;        // full-resolution disparity
  int rows, cols, Dcols, level;
  int nms_radius;        // <= 0: NMS off for this level
  int border;            // max(nonMaxSuppRadius, 3)
  float min_saliency, min_disp, max_disp;
};

__device__ __forceinline__ bool is_local_max(const float* __restrict__ S, int cols, int y, int x, int radius) {
  if (radius <= 0) return true;
  const float* p = S + (size_t) y * cols + x;
  const float v = __ldg(p);
  if (radius == 1) {
    // WITH_SIMD window: rows y-1, y+1 x cols x-1..x+2 and row y x cols x-1, x+1, x+2, strict > (masks 15/15/13)
    bool ok = (v > __ldg(p - 1)) && (v > __ldg(p + 1)) && (v > __ldg(p + 2));
    const float* u = p - cols - 1; const float* dn = p + cols - 1;
#pragma unroll
    for (int k = 0; k < 4; ++k) ok = ok && (v > __ldg(u + k)) && (v > __ldg(dn + k));
    return ok;
  }
  for (int r = -radius; r <= radius; ++r)
    for (int c = -radius; c <= radius; ++c)
      if (!(!r && !c) && __ldg(p + r * cols + c) >= v) return false;
  return true;
}

__device__ __forceinline__ bool select_pixel(const SelectArgs& a, int y, int x) {
  if (y < a.border || y >= a.rows - a.border - 1 || x < a.border || x >= a.cols - a.border - 1) return false;
  const float s = __ldg(a.S + (size_t) y * a.cols + x);
  if (!(s >= a.min_saliency)) return false;
  if (!is_local_max(a.S, a.cols, y, x, a.nms_radius)) return false;
  const float dsp = __ldg(a.D + ((size_t) 1 << a.level) * ((size_t) y * a.Dcols + x));
  return dsp >= a.min_disp && dsp <= a.max_disp;
}

constexpr int kSelPerThread = 4, kSelThreads = 256, kSelPerBlock = kSelPerThread * kSelThreads;

// pass 1: flags (u8 per pixel) and per-block counts
__global__ void __launch_bounds__(kSelThreads) select_flags_kernel(SelectArgs a, uint8_t* __restrict__ flags, int* __restrict__ block_counts) {
  __shared__ int s_warp[kSelThreads / 32];
  const int base = blockIdx.x * kSelPerBlock + threadIdx.x * kSelPerThread;
  const int total = a.rows * a.cols;
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < kSelPerThread; ++k) {
    const int p = base + k;
    uint8_t f = 0;
    if (p < total) {
      const int y = p / a.cols, x = p - y * a.cols;
      f = select_pixel(a, y, x) ? 1 : 0;
      flags[p] = f;
    }
    cnt += f;
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kSelThreads / 32; ++w) t += s_warp[w];
    block_counts[blockIdx.x] = t;
  }
}


// pass 2: exclusive scan of the block counts (single CTA), trim to a multiple of 16, shard split
__global__ void __launch_bounds__(1024) select_scan_kernel(int* __restrict__ block_counts, int nb, TemplateMeta* __restrict__ meta,
                                                            int shard_rank, int shard_size, int shard_min_points) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = (i < nb) ? block_counts[i] : 0;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = s_warp[threadIdx.x], wi = w;
      for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, wi, o); if (threadIdx.x >= o) wi += t; }
      s_warp[threadIdx.x] = wi - w;   // exclusive warp offsets
    }
    __syncthreads();
    const int excl = s_carry + s_warp[threadIdx.x >> 5] + incl - v;
    if (i < nb) block_counts[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int n_raw = s_carry;
    const int n_total = n_raw - (n_raw % 16);                 // template_data.cc:85-89
    // contiguous scan-order blocks, multiples of 16
    // ... unless the level is too small for sharding to pay (peer-memory mode): then every rank keeps everything
    const int groups = n_total / 16;
    const bool rep = shard_size > 1 && n_total < shard_min_points;
    const int g0 = rep ? 0 : (int) ((long long) groups * shard_rank / shard_size), g1 = rep ? groups : (int) ((long long) groups * (shard_rank + 1) / shard_size);
    meta->n_raw = n_raw; meta->n_total = n_total; meta->first = g0 * 16; meta->n = (g1 - g0) * 16; meta->replicated = rep ? 1 : 0;
  }
}

struct PointArgs {
  int rows, cols, Dcols, level;
  float fx, fy, cx, cy, Bf;
};

// pass 3: scatter the kept pixels in scan order; makePoint (rigid_body_warp.h:47-60, Z = Bf * (1.0/d) in double)
__global__ void __launch_bounds__(kSelThreads) select_scatter_kernel(PointArgs a, const float* __restrict__ D, const uint8_t* __restrict__ flags,
                                                                     const int* __restrict__ block_offsets, const TemplateMeta* __restrict__ meta,
                                                                     int* __restrict__ inds, float4* __restrict__ pts) {
  __shared__ int s_warp[kSelThreads / 32];
  const int base = blockIdx.x * kSelPerBlock + threadIdx.x * kSelPerThread;
  const int total = a.rows * a.cols;
  uint8_t f[kSelPerThread];
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < kSelPerThread; ++k) { f[k] = (base + k < total) ? flags[base + k] : 0; cnt += f[k]; }
  int incl = cnt;
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
  if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < (int) (threadIdx.x >> 5); ++w) woff += s_warp[w];
  int pos = block_offsets[blockIdx.x] + woff + incl - cnt;
  const int first = meta->first, n = meta->n;
#pragma unroll
  for (int k = 0; k < kSelPerThread; ++k) {
    if (!f[k]) continue;
    const int local = pos - first;
    ++pos;
    if (local < 0 || local >= n) continue;
    const int p = base + k;
    const int y = p / a.cols, x = p - y * a.cols;
    const float d = __ldg(D + ((size_t) 1 << a.level) * ((size_t) y * a.Dcols + x));
    const float Z = (float) __dmul_rn((double) a.Bf, __ddiv_rn(1.0, (double) d));
    const float X = __fmul_rn(__fmul_rn(__fsub_rn((float) x, a.cx), Z), __fdiv_rn(1.0f, a.fx));
    const float Y = __fmul_rn(__fmul_rn(__fsub_rn((float) y, a.cy), Z), __fdiv_rn(1.0f, a.fy));
    inds[local] = p;
    pts[local] = make_float4(X, Y, Z, 1.0f);
  }
}

// Hartley normalisation (warps.cc:27-48): c = mean(p), m = mean ||p - c||, s = sqrt(3)/max(m, 1e-6).
// The reference accumulates sequentially in fp32; here fp64 tree sums (deterministic order) -- c and s
// agree with the reference to ~1e-6 relative, which only rescales the (self-consistent) parametrisation.
// A small grid writes per-CTA partial sums, the last CTA (ticket) folds them into `sums[0..2]`; a one-thread
// kernel then finishes the phase.  In the point-sharded multi-GPU mode `sums` is all-reduced in between.
__global__ void __launch_bounds__(256) hartley_sum_kernel(const float4* __restrict__ pts, const TemplateMeta* __restrict__ meta,
                                                          double* __restrict__ partials, unsigned* __restrict__ ticket,
                                                          double* __restrict__ sums, int phase, int shard_rank) {
  __shared__ double s_red[8][4];
  __shared__ bool s_last;
  // a replicated level (multi-GPU) is summed by rank 0 only: the all-reduce then returns exactly the single-GPU sums
  const int n = (meta->replicated && shard_rank != 0) ? 0 : meta->n;
  double a0 = 0, a1 = 0, a2 = 0;
  const float c1 = meta->c1, c2 = meta->c2, c3 = meta->c3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    if (phase == 0) { a0 += p.x; a1 += p.y; a2 += p.z; }
    else {
      const float dx = p.x - c1, dy = p.y - c2, dz = p.z - c3;
      a0 += (double) sqrtf(dx * dx + dy * dy + dz * dz);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5][0] = a0; s_red[threadIdx.x >> 5][1] = a1; s_red[threadIdx.x >> 5][2] = a2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0, t1 = 0, t2 = 0;
    for (int w = 0; w < 8; ++w) { t0 += s_red[w][0]; t1 += s_red[w][1]; t2 += s_red[w][2]; }
    partials[blockIdx.x * 4 + 0] = t0; partials[blockIdx.x * 4 + 1] = t1; partials[blockIdx.x * 4 + 2] = t2;
    __threadfence();
    const unsigned t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {                            // fixed-order fold of the CTA partials by the whole last CTA
    __threadfence();
    double t0 = 0, t1 = 0, t2 = 0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) { t0 += __ldcg(partials + b * 4); t1 += __ldcg(partials + b * 4 + 1); t2 += __ldcg(partials + b * 4 + 2); }
    for (int o = 16; o > 0; o >>= 1) {
      t0 += __shfl_xor_sync(0xffffffffu, t0, o); t1 += __shfl_xor_sync(0xffffffffu, t1, o); t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5][0] = t0; s_red[threadIdx.x >> 5][1] = t1; s_red[threadIdx.x >> 5][2] = t2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double u0 = 0, u1 = 0, u2 = 0;
      for (int w = 0; w < 8; ++w) { u0 += s_red[w][0]; u1 += s_red[w][1]; u2 += s_red[w][2]; }
      sums[0] = u0; sums[1] = u1; sums[2] = u2;
      *ticket = 0;
    }
  }
}

__global__ void hartley_finish_kernel(TemplateMeta* __restrict__ meta, const double* __restrict__ sums, int phase) {
  const int n = meta->n_total;             // == n when unsharded
  if (phase == 0) {
    const double inv = n > 0 ? 1.0 / (double) n : 0.0;
    meta->c1 = (float) (sums[0] * inv); meta->c2 = (float) (sums[1] * inv); meta->c3 = (float) (sums[2] * inv);
  } else {
    const float m = n > 0 ? (float) (sums[0] / (double) n) : 0.0f;
    meta->s = (float) (sqrt(3.0) / (double) fmaxf(m, 1e-6f));
  }
}

__global__ void set_identity_normalization_kernel(TemplateMeta* meta) { meta->s = 1.0f; meta->c1 = meta->c2 = meta->c3 = 0.0f; }

// template records: I0 and fx*Ix, fy*Iy of every channel at every kept pixel (template_data.cc:105-131).
// One thread per (point, channel) so that consecutive lanes read consecutive floats of a 32-B pixel.
template <int C>
__global__ void __launch_bounds__(256) template_records_kernel(const float* __restrict__ desc, int cols, const int* __restrict__ inds,
                                                               const TemplateMeta* __restrict__ meta, float fx, float fy, int cd5,
                                                               float* __restrict__ gx, float* __restrict__ gy, float* __restrict__ i0) {
  const int n = meta->n;
  // grid-stride: the grid is capped (n is only known on the device)
  constexpr int S = kStride<C>;            // the padding channels of a record are written as zeros
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n * S; t += gridDim.x * blockDim.x) {
    const int i = t / S, c = t - i * S;
    if (c >= C) { i0[t] = 0.0f; gx[t] = 0.0f; gy[t] = 0.0f; continue; }
    const int p = inds[i];
    const float v0 = dval<C>(desc, p, c);
    float ix, iy;
    if (!cd5) {
      ix = __fmul_rn(0.5f, __fsub_rn(dval<C>(desc, p + 1, c), dval<C>(desc, p - 1, c)));
      iy = __fmul_rn(0.5f, __fsub_rn(dval<C>(desc, p + cols, c), dval<C>(desc, p - cols, c)));
    } else {
      const float NN = 1.0f / 18.0f;   // sic: 1/18, template_data.cc:102
      float a = __fsub_rn(__fmul_rn(1.0f, dval<C>(desc, p - 2, c)), __fmul_rn(8.0f, dval<C>(desc, p - 1, c)));
      a = __fadd_rn(a, __fmul_rn(8.0f, dval<C>(desc, p + 1, c)));
      a = __fsub_rn(a, __fmul_rn(1.0f, dval<C>(desc, p + 2, c)));
      ix = __fmul_rn(NN, a);
      float b = __fsub_rn(__fmul_rn(1.0f, dval<C>(desc, p - 2 * cols, c)), __fmul_rn(8.0f, dval<C>(desc, p - cols, c)));
      b = __fadd_rn(b, __fmul_rn(8.0f, dval<C>(desc, p + cols, c)));
      b = __fsub_rn(b, __fmul_rn(1.0f, dval<C>(desc, p + 2 * cols, c)));
      iy = __fmul_rn(NN, b);
    }
    i0[t] = v0;
    gx[t] = __fmul_rn(fx, ix);
    gy[t] = __fmul_rn(fy, iy);
  }
}

// parity dump: the reference's channel-major pixels and 1x6 Jacobians (exact-division form of
// rigid_body_warp.cc:60-315) reconstructed from the device layout
template <int C>
__global__ void __launch_bounds__(256) export_template_kernel(LevelTemplate L, float* __restrict__ pixels, float* __restrict__ J) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = L.meta->n;
  if (t >= n * C) return;
  const int i = t / C, c = t - i * C;
  const size_t rec = (size_t) i * kStride<C> + c;
  const float4 p = L.pts[i];
  const float Ix = L.gx[rec], Iy = L.gy[rec];
  const float x = p.x, y = p.y, z = p.z, z2 = __fmul_rn(z, z);
  const float xy = __fadd_rn(__fmul_rn(x, Ix), __fmul_rn(y, Iy));
  float* j = J + ((size_t) c * n + i) * 6;
  const TemplateMeta m = *L.meta;
  const float a0 = __fdiv_rn(__fmul_rn(xy, __fsub_rn(y, m.c2)), z2);
  const float t2 = __fdiv_rn(__fmul_rn(Iy, __fsub_rn(z, m.c3)), z);
  j[0] = __fsub_rn(-t2, a0);
  const float t0 = __fdiv_rn(__fmul_rn(Ix, __fsub_rn(z, m.c3)), z);
  const float t3 = __fdiv_rn(__fmul_rn(xy, __fsub_rn(x, m.c1)), z2);
  j[1] = __fadd_rn(t0, t3);
  j[2] = __fdiv_rn(__fsub_rn(__fmul_rn(Iy, __fsub_rn(x, m.c1)), __fmul_rn(Ix, __fsub_rn(y, m.c2))), z);
  const float zs = __fmul_rn(z, m.s);
  j[3] = __fdiv_rn(Ix, zs);
  j[4] = __fdiv_rn(Iy, zs);
  const float s_i = (float) (1.0 / (double) m.s);
  j[5] = -__fdiv_rn(__fmul_rn(s_i, xy), z2);
  pixels[(size_t) c * n + i] = L.i0[rec];
}

}  // namespace bp
