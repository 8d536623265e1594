// stereo.cu -- the disparity producer in front of the path, on the GPU.
//
// Replaces bpvo::StereoAlgorithm (utils/stereo_algorithm.h:14-37) for its default algorithm, "BlockMatching"
// (utils/stereo_algorithm.cc:67-85 configures OpenCV's CvStereoBMState, :99-111 runs cvFindStereoCorrespondenceBM and converts
// CV_16S -> CV_32F with 1/16).  The arithmetic is OpenCV's block matcher, integer throughout, reproduced BIT-EXACTLY:
//   k_bm_prefilter   XSOBEL pre-filter: clip(Sobel_x, -cap, cap) + cap; rows reflected; border columns and the unpaired last row = cap
//   k_bm_match       per pixel of the valid region: SAD over the w x w window for every disparity, first minimum, texture and
//                    uniqueness tests, parabola-like sub-pixel step in 1/16 px, filtered value (minDisparity - 1) elsewhere
// (definitions: oracle/stereo_oracle.cc header; parity: tests/test_gpu_stereo.py against the oracle and cv2-4.13 golden vectors).
//
// Layout: one CTA = a 32-column x 32-row tile, blockDim = (32 columns, ndisp / 16 disparity groups); the pre-filtered left / right
// tiles (+ window halo, + ndisp - 1 columns of the right image) are staged in shared memory once.  A thread owns one column and 16
// disparities: it marches down the rows keeping the 16 window SADs in registers (add the entering row's horizontal sums,
// subtract the leaving row's), and the per-pixel decisions (arg-min over all groups, uniqueness, the two neighbours of the
// minimum) go through small shared-memory arrays, double-buffered by row parity so that a row costs two CTA barriers.
// k_bm_match_fast<NW> (minDisparity == 0, the reference's setting) packs four window columns per VABSDIFF4.U8.ACC instruction;
// k_bm_match (any minDisparity <= 0, where the right-image column is clamped per window column) works byte by byte.
// No atomics, no global scratch: traffic = the two u8 images in, the disparity map out.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <string>

#include "engine_internal.h"

namespace {

#define ST_TRY(expr)                                                                           \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return bp_fail(BPVO_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

constexpr int kTX = 32, kDG = 16;                // tile columns, disparities per thread
constexpr int kMaxTileRows = 32;
constexpr int kMaxWsz = 31, kMaxDisp = 256;

__global__ void __launch_bounds__(256) k_bm_prefilter(const uint8_t* __restrict__ src0, const uint8_t* __restrict__ src1,
                                                      uint8_t* __restrict__ dst0, uint8_t* __restrict__ dst1, int rows, int cols, int cap) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= cols) return;
  const uint8_t* src = blockIdx.z ? src1 : src0;
  uint8_t* dst = blockIdx.z ? dst1 : dst0;
  int v = cap;
  if (rows >= 2 && y < (rows & ~1) && x > 0 && x < cols - 1) {        // the filter works on row pairs: an unpaired last row stays at `cap`
    const int ym = (y > 0) ? y - 1 : y + 1, yp = (y < rows - 1) ? y + 1 : y - 1;
    const uint8_t *r0 = src + (size_t) ym * cols + x, *r1 = src + (size_t) y * cols + x, *r2 = src + (size_t) yp * cols + x;
    const int s = ((int) r0[1] - (int) r0[-1]) + 2 * ((int) r1[1] - (int) r1[-1]) + ((int) r2[1] - (int) r2[-1]);
    v = min(max(s, -cap), cap) + cap;
  }
  dst[(size_t) y * cols + x] = (uint8_t) v;
}

__global__ void __launch_bounds__(256) k_bm_fill(int16_t* __restrict__ d16, float* __restrict__ df, size_t n, int16_t v16, float vf) {
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    if (d16) d16[i] = v16;
    if (df) df[i] = vf;
  }
}

struct BmArgs {
  const uint8_t* PL; const uint8_t* PR;
  int rows, cols;
  int ndisp, wsz, mindisp, cap, tex, uniq;
  int xmin, xmax, ymin, ymax;      // valid region (xmax already cut at cols + minDisparity)
  int m;                           // lofs - rofs = ndisp - 1 + minDisparity: right column = left column - m + d
  int ty;                          // output rows per tile
  int LWp, RWp;                    // shared-memory row strides of the left / right tile (multiples of 4)
  int16_t* d16; float* df;         // either may be null
};

// |a.b0 - b.b0| + ... + |a.b3 - b.b3| + c in ONE instruction (VABSDIFF4.U8.ACC)
__device__ __forceinline__ unsigned vsad4(unsigned a, unsigned b, unsigned c) {
  unsigned d;
  asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

struct BmShared { uint8_t* Lt; uint8_t* Rt; unsigned* keys; int* sads; int* flag; int sstride; };

// shared-memory plan (host and device): [left tile][right tile][keys 2 x ng x 32][SADs 32 x (ndisp + 4)][flags 2 x 32]
__host__ __device__ inline size_t bm_smem_plan(int ty, int wsz, int LWp, int RWp, int ndisp, size_t off[5]) {
  size_t o = 0;
  off[0] = o; o += (size_t) (ty + wsz - 1) * LWp;
  off[1] = o; o += (size_t) (ty + wsz - 1) * RWp;
  o = (o + 15) & ~(size_t) 15;
  off[2] = o; o += (size_t) 2 * (ndisp / kDG) * kTX * 4;
  off[3] = o; o += (size_t) kTX * (ndisp + 4) * 4;
  off[4] = o; o += (size_t) 2 * kTX * 4;
  return o;
}

// stage the pre-filtered tiles: one warp per row, no divisions; columns past the image repeat the last one (never used by a
// valid pixel)
__device__ __forceinline__ void bm_stage(const BmArgs& a, const BmShared& s, int X0, int y0, int trows, int lane, int g, int ng) {
  const int w2 = a.wsz / 2, rc0 = X0 - a.m - w2;
  for (int r = g; r < trows; r += ng) {
    const uint8_t* srcL = a.PL + (size_t) (y0 - w2 + r) * a.cols;
    const uint8_t* srcR = a.PR + (size_t) (y0 - w2 + r) * a.cols;
    for (int c = lane; c < a.LWp; c += kTX) s.Lt[r * a.LWp + c] = srcL[min(X0 - w2 + c, a.cols - 1)];
    for (int c = lane; c < a.RWp; c += kTX) s.Rt[r * a.RWp + c] = srcR[min(max(rc0 + c, 0), a.cols - 1)];
  }
  if (g == 0) { s.flag[lane] = 0; s.flag[kTX + lane] = 0; }
}

// The per-pixel decisions of one output row, given this thread's 16 window SADs: first minimum over all disparity groups,
// texture and uniqueness tests, sub-pixel step.  Two CTA barriers; keys / flags are double-buffered by row parity.
__device__ __forceinline__ void bm_decide(const BmArgs& a, const BmShared& s, const int (&sad)[kDG], int tsum, int buf, int lane, int g, int ng,
                                          int X, int y) {
  unsigned key = 0xffffffffu;                              // (SAD << 8) | internal d: the minimum is the FIRST smallest SAD
#pragma unroll
  for (int k = 0; k < kDG; ++k) key = min(key, ((unsigned) sad[k] << 8) | (unsigned) (g * kDG + k));
  s.keys[(buf * ng + g) * kTX + lane] = key;
  __syncthreads();
  // every SAD of the pixel goes to shared memory for the two neighbours of the minimum: written BETWEEN the two barriers (the
  // previous row's reader, a g == 0 thread, is past its reads once it has arrived at the barrier above)
  int4* mine = reinterpret_cast<int4*>(s.sads + lane * s.sstride + g * kDG);          // row stride ndisp + 4 ints: conflict-free 128-bit stores
#pragma unroll
  for (int q = 0; q < kDG / 4; ++q) mine[q] = make_int4(sad[4 * q], sad[4 * q + 1], sad[4 * q + 2], sad[4 * q + 3]);
  unsigned best = 0xffffffffu;
  for (int q = 0; q < ng; ++q) best = min(best, s.keys[(buf * ng + q) * kTX + lane]);
  const int mind = (int) (best & 255u), minsad = (int) (best >> 8);
  if (a.uniq > 0) {
    const int thresh = minsad + (minsad * a.uniq / 100);
    bool viol = false;
#pragma unroll
    for (int k = 0; k < kDG; ++k) viol = viol || ((unsigned) (g * kDG + k - (mind - 1)) > 2u && sad[k] <= thresh);       // d outside [mind-1, mind+1]
    if (viol) s.flag[buf * kTX + lane] = 1;
  }
  if (g == 0) s.flag[(buf ^ 1) * kTX + lane] = 0;          // the other buffer: read by the previous row's last step, set again by the next row
  __syncthreads();
  if (g == 0 && X < a.xmax) {
    int16_t out = (int16_t) ((a.mindisp - 1) * 16);
    if (tsum >= a.tex && !s.flag[buf * kTX + lane]) {
      const int dp = (mind + 1 < a.ndisp) ? mind + 1 : a.ndisp - 2, dn = (mind > 0) ? mind - 1 : 1;      // neighbours, mirrored at the ends
      const int p = s.sads[lane * s.sstride + dp], n = s.sads[lane * s.sstride + dn];
      const int den = p + n - 2 * minsad + abs(p - n);
      out = (int16_t) (((a.ndisp - mind - 1 + a.mindisp) * 256 + (den != 0 ? (p - n) * 256 / den : 0) + 15) >> 4);
    }
    const size_t o = (size_t) y * a.cols + X;
    if (a.d16) a.d16[o] = out;
    if (a.df) a.df[o] = (float) out * 0.0625f;
  }
}

__device__ __forceinline__ BmShared bm_shared(const BmArgs& a, unsigned char* smem) {
  size_t off[5];
  bm_smem_plan(a.ty, a.wsz, a.LWp, a.RWp, a.ndisp, off);
  BmShared s;
  s.Lt = smem + off[0]; s.Rt = smem + off[1]; s.keys = reinterpret_cast<unsigned*>(smem + off[2]);
  s.sads = reinterpret_cast<int*>(smem + off[3]); s.flag = reinterpret_cast<int*>(smem + off[4]); s.sstride = a.ndisp + 4;
  return s;
}

// ---- generic matcher: any window, any minDisparity <= 0 (the right-image column is clamped per window column) ----
template <int SIGN>
__device__ __forceinline__ void bm_row(const uint8_t* __restrict__ lrow, const uint8_t* __restrict__ rrow, int wsz, int rb0, int rb_max, int cap,
                                       int (&sad)[kDG], int& tsum) {
  for (int dx = 0; dx < wsz; ++dx) {
    const unsigned lv = lrow[dx];
    const uint8_t* r = rrow + min(rb0 + dx, rb_max);
    if (SIGN > 0) {
      tsum = (int) __usad(lv, (unsigned) cap, (unsigned) tsum);
#pragma unroll
      for (int k = 0; k < kDG; ++k) sad[k] = (int) __usad(lv, (unsigned) r[k], (unsigned) sad[k]);
    } else {
      tsum -= (int) __usad(lv, (unsigned) cap, 0u);
#pragma unroll
      for (int k = 0; k < kDG; ++k) sad[k] -= (int) __usad(lv, (unsigned) r[k], 0u);
    }
  }
}

__global__ void __launch_bounds__(kTX * (kMaxDisp / kDG)) k_bm_match(const BmArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x, g = threadIdx.y, ng = blockDim.y;
  const int w2 = a.wsz / 2;
  const int X0 = a.xmin + blockIdx.x * kTX, y0 = a.ymin + blockIdx.y * a.ty;
  const int ty = min(a.ty, a.ymax - y0);                  // output rows of this tile
  const int trows = ty + a.wsz - 1;                       // staged rows: y0 - w2 .. y0 + ty - 1 + w2 (all inside the image)
  const int rc0 = X0 - a.m - w2;                          // first staged column of the right image
  const BmShared s = bm_shared(a, smem);
  bm_stage(a, s, X0, y0, trows, lane, g, ng);
  __syncthreads();
  const int X = X0 + lane;
  // right-image base column of window column dx: min(X - m - w2 + dx, cols - ndisp) (+ d), relative to the staged tile
  const int rb0 = lane + g * kDG, rb_max = a.cols - a.ndisp - rc0 + g * kDG;
  int sad[kDG], tsum = 0;
#pragma unroll
  for (int k = 0; k < kDG; ++k) sad[k] = 0;
  for (int r = 0; r < a.wsz; ++r) bm_row<1>(s.Lt + (size_t) r * a.LWp + lane, s.Rt + (size_t) r * a.RWp, a.wsz, rb0, rb_max, a.cap, sad, tsum);
  for (int yy = 0; yy < ty; ++yy) {
    if (yy > 0) {
      bm_row<1>(s.Lt + (size_t) (yy + a.wsz - 1) * a.LWp + lane, s.Rt + (size_t) (yy + a.wsz - 1) * a.RWp, a.wsz, rb0, rb_max, a.cap, sad, tsum);
      bm_row<-1>(s.Lt + (size_t) (yy - 1) * a.LWp + lane, s.Rt + (size_t) (yy - 1) * a.RWp, a.wsz, rb0, rb_max, a.cap, sad, tsum);
    }
    bm_decide(a, s, sad, tsum, yy & 1, lane, g, ng, X, y0 + yy);
  }
}

// ---- fast matcher (minDisparity == 0: no clamping inside the valid region): four window columns per instruction ----
// A window row of the left image is NW packed words (bytes past the window masked to zero); for disparity index k the matching
// right-image bytes are the same row shifted by k bytes: every 4-byte window at every byte offset of a (16 + 4 NW)-byte
// register strip, cut out with funnel shifts.  One VABSDIFF4.U8.ACC then adds four absolute differences to a SAD:
// per thread and tile row NW + 1 + NW + 5 aligned 32-bit shared-memory loads and ~22 NW + 20 instructions for 16 disparities
// (the byte-wise version: 17 loads and ~35 instructions per window COLUMN).  Entering and leaving rows accumulate separately
// (the instruction only adds); SAD = entered - left.
template <int NW>
__device__ __forceinline__ void bm_row_fast(const uint8_t* __restrict__ lrow_w, const uint8_t* __restrict__ rrow_w, unsigned sh, unsigned tail,
                                            unsigned capw, unsigned (&acc)[kDG], unsigned& tacc) {
  const unsigned* lw = reinterpret_cast<const unsigned*>(lrow_w);
  const unsigned* rw = reinterpret_cast<const unsigned*>(rrow_w);
  unsigned La[NW + 1], Ra[NW + 5];
#pragma unroll
  for (int i = 0; i < NW + 1; ++i) La[i] = lw[i];
#pragma unroll
  for (int i = 0; i < NW + 5; ++i) Ra[i] = rw[i];
  unsigned Lw[NW], Bw[NW + 4];
#pragma unroll
  for (int i = 0; i < NW; ++i) Lw[i] = __funnelshift_r(La[i], La[i + 1], sh);
  Lw[NW - 1] &= tail;
#pragma unroll
  for (int i = 0; i < NW + 4; ++i) Bw[i] = __funnelshift_r(Ra[i], Ra[i + 1], sh);
#pragma unroll
  for (int j = 0; j < NW; ++j) tacc = vsad4(Lw[j], (j == NW - 1) ? (capw & tail) : capw, tacc);
#pragma unroll
  for (int k = 0; k < kDG; ++k) {
#pragma unroll
    for (int j = 0; j < NW; ++j) {
      const int o = k + 4 * j;
      unsigned W = (o % 4 == 0) ? Bw[o / 4] : __funnelshift_r(Bw[o / 4], Bw[o / 4 + 1], 8 * (o % 4));
      if (j == NW - 1) W &= tail;
      acc[k] = vsad4(W, Lw[j], acc[k]);
    }
  }
}

template <int NW>
__global__ void __launch_bounds__(kTX * (kMaxDisp / kDG)) k_bm_match_fast(const BmArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x, g = threadIdx.y, ng = blockDim.y;
  const int X0 = a.xmin + blockIdx.x * kTX, y0 = a.ymin + blockIdx.y * a.ty;
  const int ty = min(a.ty, a.ymax - y0);
  const int trows = ty + a.wsz - 1;
  const BmShared s = bm_shared(a, smem);
  bm_stage(a, s, X0, y0, trows, lane, g, ng);
  __syncthreads();
  const int X = X0 + lane;
  const unsigned sh = 8u * (unsigned) (lane & 3);                                  // byte offset of this thread's strips inside their aligned words
  const unsigned tail = 0xffffffffu >> (8 * (4 * NW - a.wsz));                    // keeps the window bytes of the last word
  const unsigned capw = (unsigned) a.cap * 0x01010101u;
  const uint8_t* lbase = s.Lt + (lane & ~3);
  const uint8_t* rbase = s.Rt + (lane & ~3) + g * kDG;
  unsigned accp[kDG], accm[kDG], tp = 0, tm = 0;
#pragma unroll
  for (int k = 0; k < kDG; ++k) { accp[k] = 0; accm[k] = 0; }
  for (int r = 0; r < a.wsz; ++r) bm_row_fast<NW>(lbase + (size_t) r * a.LWp, rbase + (size_t) r * a.RWp, sh, tail, capw, accp, tp);
  for (int yy = 0; yy < ty; ++yy) {
    if (yy > 0) {
      bm_row_fast<NW>(lbase + (size_t) (yy + a.wsz - 1) * a.LWp, rbase + (size_t) (yy + a.wsz - 1) * a.RWp, sh, tail, capw, accp, tp);
      bm_row_fast<NW>(lbase + (size_t) (yy - 1) * a.LWp, rbase + (size_t) (yy - 1) * a.RWp, sh, tail, capw, accm, tm);
    }
    int sad[kDG];
#pragma unroll
    for (int k = 0; k < kDG; ++k) sad[k] = (int) (accp[k] - accm[k]);
    bm_decide(a, s, sad, (int) (tp - tm), yy & 1, lane, g, ng, X, y0 + yy);
  }
}

typedef void (*BmKernel)(const BmArgs);
static BmKernel pick_fast(int nw) {
  switch (nw) {
    case 2: return k_bm_match_fast<2>; case 3: return k_bm_match_fast<3>; case 4: return k_bm_match_fast<4>; case 5: return k_bm_match_fast<5>;
    case 6: return k_bm_match_fast<6>; case 7: return k_bm_match_fast<7>; default: return k_bm_match_fast<8>;
  }
}

static bool device_readable(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

}  // namespace

struct bpvo_b200_stereo {
  bpvo_b200_stereo_params p;
  int rows = 0, cols = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  uint8_t* d_in[2] = {}; uint8_t* d_pre[2] = {};
  int16_t* d_d16 = nullptr; float* d_df = nullptr;
  size_t smem = 0;
  int ty = 32, LWp = 0, RWp = 0;
  BmKernel kernel = nullptr;
  long long launches = 0;
};

extern "C" {

void bpvo_b200_stereo_default_params(bpvo_b200_stereo_params* p) {        // utils/stereo_algorithm.cc:67-85 (OpenCV's own defaults)
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->numberOfDisparities = 0;        // "must be provided" (:75)
  p->SADWindowSize = 15; p->minDisparity = 0;
  p->preFilterType = BPVO_B200_STEREO_BM_XSOBEL; p->preFilterSize = 9; p->preFilterCap = 31;
  p->textureThreshold = 10; p->uniquenessRatio = 15; p->speckleWindowSize = 0; p->speckleRange = 0;
  p->trySmallerWindows = 0; p->disp12MaxDiff = -1; p->device_id = 0;
}

int bpvo_b200_stereo_create(bpvo_b200_stereo** out, int rows, int cols, const bpvo_b200_stereo_params* p) {
  if (!out || !p) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  // OpenCV's own argument checks (stereobm.cpp), with its messages
  if (p->preFilterType != BPVO_B200_STEREO_BM_XSOBEL && p->preFilterType != BPVO_B200_STEREO_BM_NORMALIZED_RESPONSE)
    return bp_fail(BPVO_B200_ERR_INVALID_ARG, "preFilterType must be = CV_STEREO_BM_NORMALIZED_RESPONSE or CV_STEREO_BM_XSOBEL");
  if (p->preFilterCap < 1 || p->preFilterCap > 63) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "preFilterCap must be within 1..63");
  if (p->SADWindowSize < 5 || p->SADWindowSize > 255 || p->SADWindowSize % 2 == 0 || p->SADWindowSize >= (rows < cols ? rows : cols))
    return bp_fail(BPVO_B200_ERR_INVALID_ARG, "SADWindowSize must be odd, be within 5..255 and be not larger than image width or height");
  if (p->numberOfDisparities <= 0 || p->numberOfDisparities % 16 != 0)
    return bp_fail(BPVO_B200_ERR_INVALID_ARG, "numDisparities must be positive and divisble by 16");
  if (p->textureThreshold < 0) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "texture threshold must be non-negative");
  if (p->uniquenessRatio < 0) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "uniqueness ratio must be non-negative");
  // what this engine does not accelerate
  if (p->preFilterType != BPVO_B200_STEREO_BM_XSOBEL) return bp_fail(BPVO_B200_ERR_UNSUPPORTED, "only the XSOBEL pre-filter (the reference's setting) is on the accelerated path");
  if (p->SADWindowSize > kMaxWsz) return bp_fail(BPVO_B200_ERR_UNSUPPORTED, "SADWindowSize above %d", kMaxWsz);
  if (p->numberOfDisparities > kMaxDisp) return bp_fail(BPVO_B200_ERR_UNSUPPORTED, "numberOfDisparities above %d", kMaxDisp);
  if (p->minDisparity > 0) return bp_fail(BPVO_B200_ERR_UNSUPPORTED, "minDisparity > 0 (OpenCV writes past the end of the output rows there)");
  if (p->speckleWindowSize > 0) return bp_fail(BPVO_B200_ERR_UNSUPPORTED, "speckle filtering (the reference leaves it off)");
  if (p->disp12MaxDiff >= 0) return bp_fail(BPVO_B200_ERR_UNSUPPORTED, "left-right check (the reference leaves it off)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return bp_fail(BPVO_B200_ERR_CUDA, "no CUDA device: bpvo_b200 has no CPU fallback"); }
  if (p->device_id < 0 || p->device_id >= ndev) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "device_id %d out of range (%d devices)", p->device_id, ndev);
  ST_TRY(cudaSetDevice(p->device_id));
  bpvo_b200_stereo* s = new bpvo_b200_stereo();
  s->p = *p; s->rows = rows; s->cols = cols;
  const size_t n = (size_t) rows * cols;
  cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&s->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&s->ev1);
  for (int k = 0; k < 2 && e == cudaSuccess; ++k) { e = cudaMalloc(&s->d_in[k], n); if (e == cudaSuccess) e = cudaMalloc(&s->d_pre[k], n); }
  if (e == cudaSuccess) e = cudaMalloc(&s->d_d16, n * sizeof(int16_t));
  if (e == cudaSuccess) e = cudaMalloc(&s->d_df, n * sizeof(float));
  // tile plan: 32 columns x `ty` rows per CTA; row strides cover the aligned 32-bit strips the fast matcher loads
  const int nw = (p->SADWindowSize + 3) / 4, ng = p->numberOfDisparities / kDG;
  s->LWp = kTX + 4 * nw; s->RWp = kDG * ng + 4 * nw + kTX;
  s->ty = 32;
  if (const char* ev = getenv("BPVO_B200_STEREO_TILE_ROWS")) { const int v = atoi(ev); if (v >= 1 && v <= kMaxTileRows) s->ty = v; }
  const bool generic = p->minDisparity != 0 || (getenv("BPVO_B200_STEREO_GENERIC") && atoi(getenv("BPVO_B200_STEREO_GENERIC")) != 0);
  s->kernel = generic ? k_bm_match : pick_fast(nw);
  size_t off[5];
  s->smem = bm_smem_plan(s->ty, p->SADWindowSize, s->LWp, s->RWp, p->numberOfDisparities, off);
  if (e == cudaSuccess && s->smem > 48 * 1024) e = cudaFuncSetAttribute(reinterpret_cast<const void*>(s->kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int) s->smem);
  if (e != cudaSuccess) {
    bpvo_b200_stereo_destroy(s);
    return bp_fail(BPVO_B200_ERR_CUDA, "stereo_create: %s", cudaGetErrorString(e));
  }
  *out = s;
  return BPVO_B200_OK;
}

int bpvo_b200_stereo_destroy(bpvo_b200_stereo* s) {
  if (!s) return BPVO_B200_OK;
  cudaSetDevice(s->p.device_id);
  if (s->stream) cudaStreamSynchronize(s->stream);
  for (int k = 0; k < 2; ++k) { cudaFree(s->d_in[k]); cudaFree(s->d_pre[k]); }
  cudaFree(s->d_d16); cudaFree(s->d_df);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return BPVO_B200_OK;
}

// StereoAlgorithm::run (utils/stereo_algorithm.cc:163-166).  left / right: rows x cols u8, host (pageable or pinned) or device;
// dmap (f32, disparity in pixels, invalid = minDisparity - 1) and disp16 (OpenCV's CV_16S map, 4 fractional bits): host or device,
// either may be null.  Returns after the results are in place.
int bpvo_b200_stereo_run(bpvo_b200_stereo* s, const uint8_t* left, const uint8_t* right, float* dmap, int16_t* disp16) {
  if (!s) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null stereo object");
  if (!left || !right) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "nullptr image");
  ST_TRY(cudaSetDevice(s->p.device_id));
  const int rows = s->rows, cols = s->cols;
  const size_t n = (size_t) rows * cols;
  const uint8_t* in[2] = {left, right};
  for (int k = 0; k < 2; ++k)
    if (!device_readable(in[k])) { ST_TRY(cudaMemcpyAsync(s->d_in[k], in[k], n, cudaMemcpyHostToDevice, s->stream)); in[k] = s->d_in[k]; }
  const bool df_dev = dmap && device_readable(dmap), d16_dev = disp16 && device_readable(disp16);
  float* df = df_dev ? dmap : s->d_df;
  int16_t* d16 = d16_dev ? disp16 : (disp16 ? s->d_d16 : nullptr);
  const bpvo_b200_stereo_params& p = s->p;
  ST_TRY(cudaEventRecord(s->ev0, s->stream));
  k_bm_prefilter<<<dim3((cols + 255) / 256, rows, 2), 256, 0, s->stream>>>(in[0], in[1], s->d_pre[0], s->d_pre[1], rows, cols, p.preFilterCap);
  const int16_t filtered = (int16_t) ((p.minDisparity - 1) * 16);
  k_bm_fill<<<296, 256, 0, s->stream>>>(d16, df, n, filtered, (float) filtered * 0.0625f);
  s->launches += 2;
  BmArgs a;
  a.PL = s->d_pre[0]; a.PR = s->d_pre[1]; a.rows = rows; a.cols = cols;
  a.ndisp = p.numberOfDisparities; a.wsz = p.SADWindowSize; a.mindisp = p.minDisparity; a.cap = p.preFilterCap; a.tex = p.textureThreshold; a.uniq = p.uniquenessRatio;
  const int w2 = a.wsz / 2, maxD = a.mindisp + a.ndisp - 1;
  a.m = maxD;
  a.xmin = (maxD > 0 ? maxD : 0) + w2;
  a.xmax = cols - w2 < cols + a.mindisp ? cols - w2 : cols + a.mindisp;
  a.ymin = w2; a.ymax = rows - w2;
  a.d16 = d16; a.df = df;
  a.ty = s->ty; a.LWp = s->LWp; a.RWp = s->RWp;
  const int lofs = maxD > 0 ? maxD : 0, rofs = maxD < 0 ? -maxD : 0, width1 = cols - rofs - a.ndisp + 1;
  if (a.xmax > a.xmin && a.ymax > a.ymin && lofs < cols && rofs < cols && width1 >= 1) {
    const dim3 grid((a.xmax - a.xmin + kTX - 1) / kTX, (a.ymax - a.ymin + a.ty - 1) / a.ty), block(kTX, a.ndisp / kDG);
    s->kernel<<<grid, block, s->smem, s->stream>>>(a);
    s->launches += 1;
  }
  ST_TRY(cudaGetLastError());
  ST_TRY(cudaEventRecord(s->ev1, s->stream));
  if (dmap && !df_dev) ST_TRY(cudaMemcpyAsync(dmap, s->d_df, n * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
  if (disp16 && !d16_dev) ST_TRY(cudaMemcpyAsync(disp16, s->d_d16, n * sizeof(int16_t), cudaMemcpyDeviceToHost, s->stream));
  ST_TRY(cudaStreamSynchronize(s->stream));
  return BPVO_B200_OK;
}

// StereoAlgorithm::getInvalidValue (utils/stereo_algorithm.cc:138-146, 168): short(minDisparity - 1) / 16.0f, as the reference
// computes it -- which is NOT the value invalid pixels carry in the map (the reference divides the already descaled marker by 16
// once more: -0.0625 for minDisparity 0).  The marker itself: bpvo_b200_stereo_filtered_value = minDisparity - 1.
float bpvo_b200_stereo_invalid_value(const bpvo_b200_stereo* s) { return s ? (float) (int16_t) (s->p.minDisparity - 1) / 16.0f : -0.0625f; }
float bpvo_b200_stereo_filtered_value(const bpvo_b200_stereo* s) { return s ? (float) (s->p.minDisparity - 1) : -1.0f; }

// the XSOBEL pre-filtered pair of the last run (parity dump): rows x cols u8 each, host pointers
int bpvo_b200_stereo_get_prefiltered(bpvo_b200_stereo* s, uint8_t* left, uint8_t* right) {
  if (!s) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null stereo object");
  ST_TRY(cudaSetDevice(s->p.device_id));
  const size_t n = (size_t) s->rows * s->cols;
  if (left) ST_TRY(cudaMemcpy(left, s->d_pre[0], n, cudaMemcpyDeviceToHost));
  if (right) ST_TRY(cudaMemcpy(right, s->d_pre[1], n, cudaMemcpyDeviceToHost));
  return BPVO_B200_OK;
}

// device time of the kernels of the last run (CUDA events on the object's stream), and the kernels launched so far
int bpvo_b200_stereo_last_kernel_ms(bpvo_b200_stereo* s, float* ms) {
  if (!s || !ms) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  ST_TRY(cudaEventElapsedTime(ms, s->ev0, s->ev1));
  return BPVO_B200_OK;
}
long long bpvo_b200_stereo_launches(const bpvo_b200_stereo* s) { return s ? s->launches : 0; }

// image pair in, pose out: the disparity map never leaves the device (apps feed StereoAlgorithm::run's output straight into
// VisualOdometry::addFrame, utils/dataset.cc:133 -> apps/vo_app.cc)
int bpvo_b200_vo_add_stereo_frame(bpvo_b200_vo* vo, bpvo_b200_stereo* s, const uint8_t* left, const uint8_t* right, bpvo_b200_result* result) {
  if (!vo || !s) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  const int rc = bpvo_b200_stereo_run(s, left, right, s->d_df, nullptr);
  if (rc != BPVO_B200_OK) return rc;
  // the left image is on the device already (staged by the run above unless the caller's pointer was a device pointer): addFrame
  // takes that copy, one upload per image instead of two.  Both calls return with their streams drained, so the staging buffer is
  // free again before the next pair overwrites it.
  const uint8_t* left_dev = device_readable(left) ? left : s->d_in[0];
  return bpvo_b200_vo_add_frame(vo, left_dev, s->d_df, result);
}

}  // extern "C"
