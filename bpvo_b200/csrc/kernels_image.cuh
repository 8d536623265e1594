// kernels_image.cuh -- per-frame image work: pyramid and dense descriptors (sm_100a).
//
// Replaces (reference file:line):
//   cv::pyrDown chain                       bpvo/image_pyramid.cc:43-50
//   census + 8 bit-planes + 5x5 Gaussian    bpvo/census.cc:42-91, bpvo/bitplanes_descriptor.cc:37-91
//   u8 -> f32 intensity                     bpvo/intensity_descriptor.cc:31-43
//
// All of it is HBM/L2-bound byte work (33 B per pixel for bit-planes): one pass, results written once
// in the channel-interleaved layout the alignment kernels gather from.
#pragma once

#include <cuda.h>            // CUtensorMap (the TMA descriptors are encoded on the host, engine.cu)
#include "device_types.h"

namespace bp {

// u8 pyramid levels are stored with a row pitch that is a multiple of 16 bytes: what a TMA tensor map needs
__host__ __device__ __forceinline__ int u8_pitch(int cols) { return (cols + 15) & ~15; }

__device__ __forceinline__ int reflect101(int i, int n) {
  // BORDER_REFLECT_101: ... 2 1 | 0 1 2 ... n-1 | n-2 n-3 ...
  while (i < 0 || i >= n) {
    if (n == 1) return 0;
    i = (i < 0) ? -i : 2 * (n - 1) - i;
  }
  return i;
}

// ---------------------------------------------------------------------------------------------
// pyrDown: 5x5 [1 4 6 4 1]^2 / 256, reflect-101, dst = ((C+1)/2, (R+1)/2), u8 result (sum+128)>>8.
// One thread per output pixel; rows of the 5x5 window come from L1/L2 (each source byte is read by
// ~6 threads of neighbouring lanes).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pyr_down_kernel(const uint8_t* __restrict__ src, int rows, int cols, int spitch,
                                                       uint8_t* __restrict__ dst, int drows, int dcols, int dpitch) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= dcols || y >= drows) return;
  int xs[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) xs[k] = reflect101(2 * x - 2 + k, cols);
  int acc = 0;
#pragma unroll
  for (int ky = 0; ky < 5; ++ky) {
    const uint8_t* s = src + (size_t) reflect101(2 * y - 2 + ky, rows) * spitch;
    const int row = (int) __ldg(s + xs[0]) + 4 * (int) __ldg(s + xs[1]) + 6 * (int) __ldg(s + xs[2]) +
                    4 * (int) __ldg(s + xs[3]) + (int) __ldg(s + xs[4]);
    const int wy = (ky == 0 || ky == 4) ? 1 : ((ky == 2) ? 6 : 4);
    acc += wy * row;
  }
  dst[(size_t) y * dpitch + x] = (uint8_t) ((acc + 128) >> 8);
}

// ---------------------------------------------------------------------------------------------
// intensity descriptor: f32(u8), one channel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) intensity_kernel(const uint8_t* __restrict__ src, int rows, int cols, int spitch, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * cols) { const int y = i / cols, x = i - y * cols; dst[i] = (float) src[(size_t) y * spitch + x]; }
}

// ---------------------------------------------------------------------------------------------
// cv::GaussianBlur(3x3) on CV_8U ahead of the census (census.cc:63-65, sigmaPriorToCensusTransform > 0).
// Third-party fixed-point arithmetic (OpenCV 4.x, pinned bit-exact to cv2 4.13 by the oracle's golden vectors):
// 8-bit taps [a, 256 - 2a, a], row pass in 8.8, column pass in 16.16, ONE rounding (v + 2^15) >> 16, reflect-101.
// Thread per pixel; the 3x3 u8 neighbourhood comes from L1/L2 (2 B of traffic per pixel against the 33 B of the
// descriptor that follows).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) blur3_u8_kernel(const uint8_t* __restrict__ src, int rows, int cols, int spitch, int ka, int kc,
                                                       uint8_t* __restrict__ dst /* pitch u8_pitch(cols) */) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cols || y >= rows) return;
  const int xm = reflect101(x - 1, cols), xp = reflect101(x + 1, cols);
  int h[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const uint8_t* s = src + (size_t) reflect101(y - 1 + k, rows) * spitch;
    h[k] = kc * (int) __ldg(s + x) + ka * ((int) __ldg(s + xm) + (int) __ldg(s + xp));
  }
  const int v = (kc * h[1] + ka * (h[0] + h[2]) + (1 << 15)) >> 16;
  dst[(size_t) y * u8_pitch(cols) + x] = (uint8_t) min(v, 255);
}

// ---------------------------------------------------------------------------------------------
// bit-planes descriptor, fused: census (3x3, neighbour >= centre, border rows/cols = 0) ->
// 8 bit channels -> separable 5x5 Gaussian (f32, reflect-101 on the census image) -> interleaved
// [rows][cols][8] f32 store (32 B per pixel, two 16-B stores per thread, fully coalesced).
//
// Tile: 32 x 8 output pixels per CTA of 256 threads.
//   stage 1: census bytes of the (8+4) x (32+4) halo tile -> smem            (u8 reads via L1/L2)
//   stage 2: horizontal pass of all 8 bits for (8+4) x 32 -> smem [row][bit][col]  (conflict-free)
//   stage 3: vertical pass, 8 floats per thread -> global
// Evaluation order of the taps follows OpenCV's symmetric row/column filters,
// (s0*k0 + (s-1+s1)*k1) + (s-2+s2)*k2, with separate mul/add roundings (no FMA).
// ---------------------------------------------------------------------------------------------
constexpr int kBpTW = 32, kBpTH = 8;

__device__ __forceinline__ uint8_t census_at(const uint8_t* __restrict__ img, int rows, int cols, int pitch, int y, int x) {
  if (y <= 0 || y >= rows - 1 || x <= 0 || x >= cols - 1) return 0;
  const uint8_t* p = img + (size_t) y * pitch + x;
  const uint8_t c = __ldg(p);
  unsigned v = 0;
  v |= (unsigned) (__ldg(p - pitch - 1) >= c) << 0;
  v |= (unsigned) (__ldg(p - pitch) >= c) << 1;
  v |= (unsigned) (__ldg(p - pitch + 1) >= c) << 2;
  v |= (unsigned) (__ldg(p - 1) >= c) << 3;
  v |= (unsigned) (__ldg(p + 1) >= c) << 4;
  v |= (unsigned) (__ldg(p + pitch - 1) >= c) << 5;
  v |= (unsigned) (__ldg(p + pitch) >= c) << 6;
  v |= (unsigned) (__ldg(p + pitch + 1) >= c) << 7;
  return (uint8_t) v;
}

__global__ void __launch_bounds__(256) bitplanes_kernel(const uint8_t* __restrict__ img, int rows, int cols, int pitch,
                                                        float k0, float k1, float k2, int do_blur,
                                                        float* __restrict__ out) {
  __shared__ uint8_t s_census[kBpTH + 4][kBpTW + 4];
  __shared__ float s_h[kBpTH + 4][8][kBpTW];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x0 = blockIdx.x * kBpTW, y0 = blockIdx.y * kBpTH;

  for (int i = threadIdx.x; i < (kBpTH + 4) * (kBpTW + 4); i += 256) {
    const int r = i / (kBpTW + 4), c = i % (kBpTW + 4);
    const int gy = reflect101(y0 + r - 2, rows), gx = reflect101(x0 + c - 2, cols);
    s_census[r][c] = census_at(img, rows, cols, pitch, gy, gx);
  }
  __syncthreads();

  const int gx = x0 + tx, gy = y0 + ty;
  if (!do_blur) {
    if (gx < cols && gy < rows) {
      const unsigned v = s_census[ty + 2][tx + 2];
      float4* o = reinterpret_cast<float4*>(out + ((size_t) gy * cols + gx) * 8);
      o[0] = make_float4((float) (v & 1), (float) ((v >> 1) & 1), (float) ((v >> 2) & 1), (float) ((v >> 3) & 1));
      o[1] = make_float4((float) ((v >> 4) & 1), (float) ((v >> 5) & 1), (float) ((v >> 6) & 1), (float) ((v >> 7) & 1));
    }
    return;
  }

  for (int i = threadIdx.x; i < (kBpTH + 4) * kBpTW; i += 256) {
    const int r = i / kBpTW, c = i % kBpTW;
    const unsigned m2 = s_census[r][c], m1 = s_census[r][c + 1], c0 = s_census[r][c + 2],
                   p1 = s_census[r][c + 3], p2 = s_census[r][c + 4];
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const float s0 = (float) ((c0 >> b) & 1);
      const float a = __fadd_rn((float) ((m1 >> b) & 1), (float) ((p1 >> b) & 1));
      const float bb = __fadd_rn((float) ((m2 >> b) & 1), (float) ((p2 >> b) & 1));
      float v = __fmul_rn(s0, k0);
      v = __fadd_rn(v, __fmul_rn(a, k1));
      v = __fadd_rn(v, __fmul_rn(bb, k2));
      s_h[r][b][c] = v;
    }
  }
  __syncthreads();

  if (gx < cols && gy < rows) {
    float o[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      float v = __fmul_rn(k0, s_h[ty + 2][b][tx]);
      v = __fadd_rn(v, __fmul_rn(k1, __fadd_rn(s_h[ty + 3][b][tx], s_h[ty + 1][b][tx])));
      v = __fadd_rn(v, __fmul_rn(k2, __fadd_rn(s_h[ty + 4][b][tx], s_h[ty][b][tx])));
      o[b] = v;
    }
    float4* dst = reinterpret_cast<float4*>(out + ((size_t) gy * cols + gx) * 8);
    dst[0] = make_float4(o[0], o[1], o[2], o[3]);
    dst[1] = make_float4(o[4], o[5], o[6], o[7]);
  }
}


// ---------------------------------------------------------------------------------------------
// The same descriptor with the tile traffic on the TMA engine (the default; bitplanes_kernel above stays as the A/B partner,
// BPVO_B200_NO_TMA=1):
//   in : the (8 + 6) x (64 + 6) u8 halo tile of the level image, ONE cp.async.bulk.tensor.2d into shared memory, completion on
//        an mbarrier -- the census then compares shared-memory bytes instead of issuing nine global byte loads per halo pixel.
//        The box is 96 x 14 bytes starting at column x0 - 16: the TMA wants the box's first byte 16-byte aligned in global
//        memory (a start at x0 - 3 raises "illegal instruction", scripts/micro/tma_probe.cu); out-of-image parts are zero-filled;
//   out: the 8 x 64 x 8 f32 result tile (16 KB) is assembled in shared memory and leaves as ONE cp.async.bulk.tensor.3d store
//        over the [rows][cols][8] descriptor (the TMA clips what lies outside the image: no per-thread bounds checks, full
//        128-byte write transactions).
// Arithmetic (census rule, tap order, roundings) is the other kernel's, bit for bit.
// ---------------------------------------------------------------------------------------------
constexpr int kTmTW = 64, kTmTH = 8, kTmInW = 96, kTmInX = 16, kTmInH = kTmTH + 6;     // input box: columns [x0 - 16, x0 + 80), rows [y0 - 3, y0 + 11)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {       // one lane of a converged warp (the TMA instructions want a uniform issue)
  unsigned p;
  asm volatile("{ .reg .pred P; elect.sync _|P, 0xffffffff; selp.u32 %0, 1, 0, P; }" : "=r"(p));
  return p != 0;
}

// map_in / map_out: 128-byte CUtensorMap descriptors in GLOBAL memory (written once by the host at frame creation)
__global__ void __launch_bounds__(256) bitplanes_tma_kernel(const CUtensorMap* __restrict__ map_in, const CUtensorMap* __restrict__ map_out,
                                                            int rows, int cols, float k0, float k1, float k2, int do_blur) {
  __shared__ __align__(128) uint8_t s_in[kTmInH][kTmInW];
  __shared__ __align__(128) float s_out[kTmTH][kTmTW][8];
  __shared__ uint8_t s_census[kTmTH + 4][kTmTW + 4];
  __shared__ float s_h[kTmTH + 4][8][kTmTW];
  __shared__ __align__(8) unsigned long long s_bar;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * kTmTW, y0 = blockIdx.y * kTmTH;
  const unsigned bar = smem_u32(&s_bar);

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid < 32 && elect_one()) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kTmInH * kTmInW) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(&s_in[0][0])), "l"(reinterpret_cast<unsigned long long>(map_in)), "r"(x0 - kTmInX), "r"(y0 - 3), "r"(bar) : "memory");
  }
  {
    unsigned done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar) : "memory");
  }

  // census of the (TH + 4) x (TW + 4) halo positions, reflect-101 on image coordinates, from the shared-memory tile
  for (int i = tid; i < (kTmTH + 4) * (kTmTW + 4); i += 256) {
    const int r = i / (kTmTW + 4), c = i % (kTmTW + 4);
    const int gy = reflect101(y0 + r - 2, rows), gx = reflect101(x0 + c - 2, cols);
    const int ly = gy - (y0 - 3), lx = gx - (x0 - kTmInX);
    unsigned v = 0;
    // (positions whose reflection leaves the tile only feed outputs outside the image, which the TMA store clips)
    if (gy > 0 && gy < rows - 1 && gx > 0 && gx < cols - 1 && ly >= 1 && ly < kTmInH - 1 && lx >= 1 && lx < kTmInW - 1) {
      const unsigned ce = s_in[ly][lx];
      v |= (unsigned) (s_in[ly - 1][lx - 1] >= ce) << 0;
      v |= (unsigned) (s_in[ly - 1][lx] >= ce) << 1;
      v |= (unsigned) (s_in[ly - 1][lx + 1] >= ce) << 2;
      v |= (unsigned) (s_in[ly][lx - 1] >= ce) << 3;
      v |= (unsigned) (s_in[ly][lx + 1] >= ce) << 4;
      v |= (unsigned) (s_in[ly + 1][lx - 1] >= ce) << 5;
      v |= (unsigned) (s_in[ly + 1][lx] >= ce) << 6;
      v |= (unsigned) (s_in[ly + 1][lx + 1] >= ce) << 7;
    }
    s_census[r][c] = (uint8_t) v;
  }
  __syncthreads();

  if (!do_blur) {
    for (int i = tid; i < kTmTH * kTmTW; i += 256) {
      const int py = i / kTmTW, px = i % kTmTW;
      const unsigned v = s_census[py + 2][px + 2];
      float4* o = reinterpret_cast<float4*>(&s_out[py][px][0]);
      o[0] = make_float4((float) (v & 1), (float) ((v >> 1) & 1), (float) ((v >> 2) & 1), (float) ((v >> 3) & 1));
      o[1] = make_float4((float) ((v >> 4) & 1), (float) ((v >> 5) & 1), (float) ((v >> 6) & 1), (float) ((v >> 7) & 1));
    }
  } else {
    for (int i = tid; i < (kTmTH + 4) * kTmTW; i += 256) {
      const int r = i / kTmTW, c = i % kTmTW;
      const unsigned m2 = s_census[r][c], m1 = s_census[r][c + 1], c0 = s_census[r][c + 2], p1 = s_census[r][c + 3], p2 = s_census[r][c + 4];
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const float s0 = (float) ((c0 >> b) & 1);
        const float a = __fadd_rn((float) ((m1 >> b) & 1), (float) ((p1 >> b) & 1));
        const float bb = __fadd_rn((float) ((m2 >> b) & 1), (float) ((p2 >> b) & 1));
        float v = __fmul_rn(s0, k0);
        v = __fadd_rn(v, __fmul_rn(a, k1));
        v = __fadd_rn(v, __fmul_rn(bb, k2));
        s_h[r][b][c] = v;
      }
    }
    __syncthreads();
    for (int i = tid; i < kTmTH * kTmTW; i += 256) {
      const int py = i / kTmTW, px = i % kTmTW;
      float o[8];
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        float v = __fmul_rn(k0, s_h[py + 2][b][px]);
        v = __fadd_rn(v, __fmul_rn(k1, __fadd_rn(s_h[py + 3][b][px], s_h[py + 1][b][px])));
        v = __fadd_rn(v, __fmul_rn(k2, __fadd_rn(s_h[py + 4][b][px], s_h[py][b][px])));
        o[b] = v;
      }
      float4* dst = reinterpret_cast<float4*>(&s_out[py][px][0]);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the generic-proxy writes of s_out become visible to the TMA engine
  __syncthreads();
  if (tid < 32 && elect_one()) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(reinterpret_cast<unsigned long long>(map_out)), "r"(0), "r"(x0), "r"(y0), "r"(smem_u32(&s_out[0][0][0])) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must outlive the engine's reads
  }
}

// ---------------------------------------------------------------------------------------------
// The two gradient-based descriptors (bpvo/gradient_descriptor.cc), interleaved like the others:
//   GradientDescriptor  (kIntensityAndGradient, 3 channels, stride 4): I, Ix, Iy -- gradients of the image smoothed with a
//                        Gaussian whose size OpenCV derives from sigma (cv::Size()) when sigma > 0, the intensity unsmoothed (:42-64)
//   DescriptorFields    (kDescriptorFieldsFirstOrder, 5 channels, stride 8): I, then the positive and the negative part of Ix
//                        and of Iy of the sigma1-smoothed image, each smoothed with sigma2 (imsmooth, imgproc.cc:166-171) (:78-116)
// Small multi-pass kernels over f32 planes (key-frame / per-frame work of a few MB; the alignment kernels dominate).  Arithmetic
// follows the oracle operation for operation (explicit roundings, OpenCV's symmetric tap order), so results are bit-identical.
// ---------------------------------------------------------------------------------------------
struct BlurTaps { int half; float k[17]; };      // k[0] centre tap, k[j] the taps at distance j (kernel size 2 half + 1 <= 33)

__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint8_t* __restrict__ src, int rows, int cols, int spitch, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * cols) { const int y = i / cols, x = i - y * cols; dst[i] = (float) src[(size_t) y * spitch + x]; }
}
__global__ void __launch_bounds__(256) blur_row_kernel(const float* __restrict__ src, int rows, int cols, BlurTaps t, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int y = i / cols, x = i - y * cols;
  const float* s = src + (size_t) y * cols;
  float v = __fmul_rn(__ldg(s + x), t.k[0]);
  for (int j = 1; j <= t.half; ++j) v = __fadd_rn(v, __fmul_rn(__fadd_rn(__ldg(s + reflect101(x - j, cols)), __ldg(s + reflect101(x + j, cols))), t.k[j]));
  dst[i] = v;
}
// column pass; the result goes to dst[i * dst_stride + dst_off] (a plane: stride 1, or one channel of an interleaved descriptor)
__global__ void __launch_bounds__(256) blur_col_kernel(const float* __restrict__ src, int rows, int cols, BlurTaps t, float* __restrict__ dst, int dst_stride, int dst_off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int y = i / cols, x = i - y * cols;
  float v = __fmul_rn(t.k[0], __ldg(src + i));
  for (int j = 1; j <= t.half; ++j)
    v = __fadd_rn(v, __fmul_rn(t.k[j], __fadd_rn(__ldg(src + (size_t) reflect101(y + j, rows) * cols + x), __ldg(src + (size_t) reflect101(y - j, rows) * cols + x))));
  dst[(size_t) i * dst_stride + dst_off] = v;
}
// xgradient / ygradient of imgproc.h:215-266: 0.5 (I[+1] - I[-1]), one-sided 0.5 (I1 - I0) on the first / last column (row)
__device__ __forceinline__ float grad_x(const float* __restrict__ I, int cols, int y, int x) {
  const float* s = I + (size_t) y * cols;
  const int a = (x == 0) ? 0 : x - 1, b = (x == cols - 1) ? cols - 1 : x + 1;
  return __fmul_rn(0.5f, __fsub_rn(__ldg(s + b), __ldg(s + a)));
}
__device__ __forceinline__ float grad_y(const float* __restrict__ I, int rows, int cols, int y, int x) {
  const int a = (y == 0) ? 0 : y - 1, b = (y == rows - 1) ? rows - 1 : y + 1;
  return __fmul_rn(0.5f, __fsub_rn(__ldg(I + (size_t) b * cols + x), __ldg(I + (size_t) a * cols + x)));
}
// GradientDescriptor: out[p] = {f32(u8), Ix, Iy, 0}; Is = the plane the gradients are taken from
__global__ void __launch_bounds__(256) gradient_descriptor_kernel(const uint8_t* __restrict__ img, int rows, int cols, int spitch,
                                                                  const float* __restrict__ Is, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int y = i / cols, x = i - y * cols;
  reinterpret_cast<float4*>(out)[i] = make_float4((float) img[(size_t) y * spitch + x], grad_x(Is, cols, y, x), grad_y(Is, rows, cols, y, x), 0.0f);
}
// DescriptorFields, step 1: positive / negative part of one gradient direction of Is -> two planes (splitPosNeg :78-99)
__global__ void __launch_bounds__(256) split_gradient_kernel(const float* __restrict__ Is, int rows, int cols, int dir, float* __restrict__ pos, float* __restrict__ neg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int y = i / cols, x = i - y * cols;
  const float g = dir ? grad_y(Is, rows, cols, y, x) : grad_x(Is, cols, y, x);
  pos[i] = (g >= 0.0f) ? g : 0.0f;
  neg[i] = (g < 0.0f) ? g : 0.0f;
}
// a plane into one channel of an interleaved descriptor (sigma2 <= 0: no smoothing)
__global__ void __launch_bounds__(256) plane_to_channel_kernel(const float* __restrict__ src, int n, float* __restrict__ dst, int dst_stride, int dst_off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[(size_t) i * dst_stride + dst_off] = src[i];
}
// DescriptorFields: channel 0 = f32(u8), padding channels 5..7 = 0
__global__ void __launch_bounds__(256) dfields_base_kernel(const uint8_t* __restrict__ img, int rows, int cols, int spitch, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int y = i / cols, x = i - y * cols;
  float* o = out + (size_t) i * 8;
  o[0] = (float) img[(size_t) y * spitch + x]; o[5] = 0.0f; o[6] = 0.0f; o[7] = 0.0f;
}

// interleaved [rows][cols][stride] -> planar C x rows x cols (parity dumps only)
__global__ void __launch_bounds__(256) deinterleave_kernel(const float* __restrict__ src, int n, int C, int stride, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int c = 0; c < C; ++c) dst[(size_t) c * n + i] = src[(size_t) i * stride + c];
}

}  // namespace bp
