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
// k_bm_match layout: one CTA = a 32-column x 32-row tile, blockDim = (32 columns, ndisp / 16 disparity groups); the pre-filtered
// left / right tiles (+ window halo, + ndisp - 1 columns of the right image) are staged in shared memory once.  A thread owns one
// column and 16 disparities: it marches down the rows keeping the 16 window SADs in registers (add the entering row's horizontal
// sums, subtract the leaving row's), and the per-pixel decisions (arg-min over all groups, uniqueness, the two neighbours of the
// minimum) go through three small shared-memory arrays, double-buffered so that a row costs two CTA barriers.
// No atomics, no global scratch: traffic = the two u8 images in, the disparity map out.
#include <cuda_runtime.h>
#include <stdint.h>
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

constexpr int kTX = 32, kTY = 32, kDG = 16;      // tile columns / rows, disparities per thread
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
  int16_t* d16; float* df;         // either may be null
};

// horizontal window sums of one tile row into the 16 running SADs (SIGN = +1 entering row, -1 leaving row)
template <int SIGN>
__device__ __forceinline__ void bm_row(const uint8_t* __restrict__ lrow, const uint8_t* __restrict__ rrow, int wsz, int rb0, int rb_max, int cap,
                                       int (&sad)[kDG], int& tsum) {
  for (int dx = 0; dx < wsz; ++dx) {
    const int lv = lrow[dx];
    tsum += SIGN * abs(lv - cap);
    const uint8_t* r = rrow + min(rb0 + dx, rb_max);
#pragma unroll
    for (int k = 0; k < kDG; ++k) sad[k] += SIGN * abs(lv - (int) r[k]);
  }
}

__global__ void __launch_bounds__(kTX * (kMaxDisp / kDG)) k_bm_match(const BmArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x, g = threadIdx.y, ng = blockDim.y, tid = g * kTX + lane, nthreads = ng * kTX;
  const int w2 = a.wsz / 2;
  const int X0 = a.xmin + blockIdx.x * kTX, y0 = a.ymin + blockIdx.y * kTY;
  const int ty = min(kTY, a.ymax - y0);                   // output rows of this tile
  const int trows = ty + a.wsz - 1;                       // staged rows: y0 - w2 .. y0 + ty - 1 + w2 (all inside the image)
  const int LW = kTX + a.wsz - 1, RW = kTX + a.wsz - 1 + a.ndisp - 1;
  const int rc0 = X0 - a.m - w2;                          // first staged column of the right image (>= 0)
  uint8_t* Lt = smem;
  uint8_t* Rt = Lt + (size_t) (kTY + a.wsz - 1) * LW;
  unsigned* keys = reinterpret_cast<unsigned*>(smem + (((size_t) (kTY + a.wsz - 1) * (LW + RW) + 15) & ~(size_t) 15));        // [2][ng][32]
  int* pn = reinterpret_cast<int*>(keys + 2 * ng * kTX);                                                                      // [2][2][32]
  int* flag = pn + 2 * 2 * kTX;                                                                                               // [2][32]
  for (int i = tid; i < trows * LW; i += nthreads) {
    const int r = i / LW, c = i - r * LW;
    Lt[i] = a.PL[(size_t) (y0 - w2 + r) * a.cols + min(X0 - w2 + c, a.cols - 1)];
  }
  for (int i = tid; i < trows * RW; i += nthreads) {
    const int r = i / RW, c = i - r * RW;
    Rt[i] = a.PR[(size_t) (y0 - w2 + r) * a.cols + min(rc0 + c, a.cols - 1)];
  }
  if (tid < 2 * kTX) flag[tid] = 0;
  __syncthreads();

  const int X = X0 + lane;
  // right-image base column of window column dx: min(X - m - w2 + dx, cols - ndisp) (+ d), relative to the staged tile
  const int rb0 = X - a.m - w2 - rc0 + g * kDG, rb_max = a.cols - a.ndisp - rc0 + g * kDG;
  int sad[kDG], tsum = 0;
#pragma unroll
  for (int k = 0; k < kDG; ++k) sad[k] = 0;
  for (int r = 0; r < a.wsz; ++r) bm_row<1>(Lt + (size_t) r * LW + lane, Rt + (size_t) r * RW, a.wsz, rb0, rb_max, a.cap, sad, tsum);
  const int16_t filtered = (int16_t) ((a.mindisp - 1) * 16);
  for (int yy = 0; yy < ty; ++yy) {
    if (yy > 0) {
      bm_row<1>(Lt + (size_t) (yy + a.wsz - 1) * LW + lane, Rt + (size_t) (yy + a.wsz - 1) * RW, a.wsz, rb0, rb_max, a.cap, sad, tsum);
      bm_row<-1>(Lt + (size_t) (yy - 1) * LW + lane, Rt + (size_t) (yy - 1) * RW, a.wsz, rb0, rb_max, a.cap, sad, tsum);
    }
    const int buf = yy & 1;
    unsigned key = 0xffffffffu;                            // (SAD << 8) | internal d: the minimum is the FIRST smallest SAD
#pragma unroll
    for (int k = 0; k < kDG; ++k) key = min(key, ((unsigned) sad[k] << 8) | (unsigned) (g * kDG + k));
    keys[(buf * ng + g) * kTX + lane] = key;
    __syncthreads();
    unsigned best = 0xffffffffu;
    for (int q = 0; q < ng; ++q) best = min(best, keys[(buf * ng + q) * kTX + lane]);
    const int mind = (int) (best & 255u), minsad = (int) (best >> 8);
    const int thresh = minsad + (minsad * a.uniq / 100);
    const int dp = (mind + 1 < a.ndisp) ? mind + 1 : a.ndisp - 2, dn = (mind > 0) ? mind - 1 : 1;      // neighbours, mirrored at the ends
    bool viol = false;
#pragma unroll
    for (int k = 0; k < kDG; ++k) {
      const int d = g * kDG + k;
      viol = viol || ((d < mind - 1 || d > mind + 1) && sad[k] <= thresh);
      if (d == dp) pn[(buf * 2 + 0) * kTX + lane] = sad[k];
      if (d == dn) pn[(buf * 2 + 1) * kTX + lane] = sad[k];
    }
    if (a.uniq > 0 && viol) flag[buf * kTX + lane] = 1;
    if (g == 0) flag[(buf ^ 1) * kTX + lane] = 0;          // the other buffer: read by the previous row's last step, set again by the next row
    __syncthreads();
    if (g == 0 && X < a.xmax) {
      int16_t out = filtered;
      if (tsum >= a.tex && !flag[buf * kTX + lane]) {
        const int p = pn[(buf * 2 + 0) * kTX + lane], n = pn[(buf * 2 + 1) * kTX + lane];
        const int den = p + n - 2 * minsad + abs(p - n);
        out = (int16_t) (((a.ndisp - mind - 1 + a.mindisp) * 256 + (den != 0 ? (p - n) * 256 / den : 0) + 15) >> 4);
      }
      const size_t o = (size_t) (y0 + yy) * a.cols + X;
      if (a.d16) a.d16[o] = out;
      if (a.df) a.df[o] = (float) out * 0.0625f;
    }
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
  const int LW = kTX + p->SADWindowSize - 1, RW = LW + p->numberOfDisparities - 1, ng = p->numberOfDisparities / kDG;
  s->smem = (((size_t) (kTY + p->SADWindowSize - 1) * (LW + RW) + 15) & ~(size_t) 15) + (size_t) (2 * ng * kTX + 2 * 2 * kTX + 2 * kTX) * 4;
  if (e == cudaSuccess && s->smem > 48 * 1024) e = cudaFuncSetAttribute(k_bm_match, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) s->smem);
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
  const int lofs = maxD > 0 ? maxD : 0, rofs = maxD < 0 ? -maxD : 0, width1 = cols - rofs - a.ndisp + 1;
  if (a.xmax > a.xmin && a.ymax > a.ymin && lofs < cols && rofs < cols && width1 >= 1) {
    const dim3 grid((a.xmax - a.xmin + kTX - 1) / kTX, (a.ymax - a.ymin + kTY - 1) / kTY), block(kTX, a.ndisp / kDG);
    k_bm_match<<<grid, block, s->smem, s->stream>>>(a);
    s->launches += 1;
  }
  ST_TRY(cudaGetLastError());
  ST_TRY(cudaEventRecord(s->ev1, s->stream));
  if (dmap && !df_dev) ST_TRY(cudaMemcpyAsync(dmap, s->d_df, n * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
  if (disp16 && !d16_dev) ST_TRY(cudaMemcpyAsync(disp16, s->d_d16, n * sizeof(int16_t), cudaMemcpyDeviceToHost, s->stream));
  ST_TRY(cudaStreamSynchronize(s->stream));
  return BPVO_B200_OK;
}

// StereoAlgorithm::getInvalidValue (utils/stereo_algorithm.cc:138-146, 168)
float bpvo_b200_stereo_invalid_value(const bpvo_b200_stereo* s) { return s ? (float) (int16_t) (s->p.minDisparity - 1) * 16.0f / 16.0f : -1.0f; }

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
  return bpvo_b200_vo_add_frame(vo, left, s->d_df, result);
}

}  // extern "C"
