// engine.cu -- host side of the seam-level C ABI (include/bpvo_b200.h): device buffers, stream,
// kernel launches, the host-driven and the on-device Gauss-Newton drivers.
//
// Mirrors, method for method, what bpvo/vo.cc calls on VisualOdometryFrame (bpvo/vo_frame.{h,cc}) and
// VisualOdometryPoseEstimator (bpvo/vo_pose_estimator.{h,cc}); the state machine of vo.cc itself
// lives in host/vo_shim.cpp on top of these entry points.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/bpvo_b200.h"
#include "device_types.h"
#include "kernels_image.cuh"
#include "kernels_template.cuh"
#include "kernels_linearize.cuh"
#include "engine_internal.h"

using namespace bp;

namespace {
thread_local std::string g_err;
}

int bp_fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  g_err = buf;
  return code;
}

cudaError_t bp_join_copies(bpvo_b200_ctx* c) {
  if (!c->disp_pending) return cudaSuccess;
  c->disp_pending = false;
  return cudaStreamWaitEvent(c->stream, c->disp_done, 0);
}
cudaError_t bp_sync_stream(bpvo_b200_ctx* c) {
  const cudaError_t e = bp_join_copies(c);
  return e != cudaSuccess ? e : cudaStreamSynchronize(c->stream);
}

#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return bp_fail(BPVO_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define LAUNCH_CHECK(ctx)                                                                      \
  do {                                                                                         \
    (ctx)->counters.launches++;                                                                \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess)                                                                     \
      return bp_fail(BPVO_B200_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// channel-count dispatch: CC is a compile-time constant inside the statement (1 intensity, 3 intensity + gradient,
// 5 descriptor fields, 8 bit-planes)
#define BP_SWITCH_C(CH, ...)                                     \
  switch (CH) {                                                  \
    case 1: { constexpr int CC = 1; __VA_ARGS__; } break;        \
    case 3: { constexpr int CC = 3; __VA_ARGS__; } break;        \
    case 5: { constexpr int CC = 5; __VA_ARGS__; } break;        \
    default: { constexpr int CC = 8; __VA_ARGS__; } break;       \
  }

// -------------------------------------------------------------------------------------------------
// phase timers (cudaEvent pairs, only when profiling is on)
// -------------------------------------------------------------------------------------------------
struct PhaseTimer {
  bpvo_b200_ctx* c; double* acc; bool on;
  PhaseTimer(bpvo_b200_ctx* ctx, double* a) : c(ctx), acc(a), on(ctx->profiling) {
    if (on) cudaEventRecord(c->ev0, c->stream);
  }
  ~PhaseTimer() {
    if (on) {
      cudaEventRecord(c->ev1, c->stream);
      cudaEventSynchronize(c->ev1);
      float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1);
      *acc += ms; c->counters.ms_total += ms;
    }
  }
};

static int ctx_init(bpvo_b200_ctx* c, const float K[9], float baseline, int rows, int cols, const bpvo_b200_params* p);
static int frame_init(bpvo_b200_ctx* c, bpvo_b200_frame* f);

// TMA descriptors of a frame's level images (u8, pitched) and descriptors (f32 [rows][cols][8]) for bitplanes_tma_kernel.
// cuTensorMapEncodeTiled is a driver entry point: fetched through the runtime, no link against libcuda.
static bool encode_tensor_maps(bpvo_b200_ctx* c, bpvo_b200_frame* f) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess || !fn) {
    cudaGetLastError();
    return false;
  }
  EncodeFn encode = reinterpret_cast<EncodeFn>(fn);
  CUtensorMap maps[2 * kMaxLevels];
  memset(maps, 0, sizeof(maps));
  for (int l = c->p.maxTestLevel; l < c->L; ++l) {
    const LevelGeom& g = c->geom[l];
    {
      const cuuint64_t dims[2] = {(cuuint64_t) g.cols, (cuuint64_t) g.rows};
      const cuuint64_t strides[1] = {(cuuint64_t) u8_pitch(g.cols)};
      const cuuint32_t box[2] = {(cuuint32_t) kTmInW, (cuuint32_t) kTmInH}, es[2] = {1, 1};
      if (encode(&maps[2 * l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, f->pyr[l], dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    }
    {
      const cuuint64_t dims[3] = {8, (cuuint64_t) g.cols, (cuuint64_t) g.rows};
      const cuuint64_t strides[2] = {32, (cuuint64_t) g.cols * 32};
      const cuuint32_t box[3] = {8, (cuuint32_t) kTmTW, (cuuint32_t) kTmTH}, es[3] = {1, 1, 1};
      if (encode(&maps[2 * l + 1], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, f->desc[l], dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    }
  }
  // the descriptors live in global memory ([level][in, out], 64-byte aligned): written once, before any kernel reads them
  if (cudaMalloc(&f->d_maps, sizeof(maps)) != cudaSuccess) { cudaGetLastError(); f->d_maps = nullptr; return false; }
  if (cudaMemcpyAsync(f->d_maps, maps, sizeof(maps), cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
      bp_sync_stream(c) != cudaSuccess) { cudaGetLastError(); return false; }
  return true;
}

extern "C" {

int bpvo_b200_version(void) { return BPVO_B200_VERSION; }
const char* bpvo_b200_last_error(void) { return g_err.c_str(); }

void bpvo_b200_default_params(bpvo_b200_params* p) {     // bpvo/types.cc:31-66
  p->numPyramidLevels = -1; p->minImageDimensionForPyramid = 40;
  p->sigmaPriorToCensusTransform = -1.0f; p->sigmaBitPlanes = 0.5f;
  p->maxIterations = 50; p->parameterTolerance = 1e-7f; p->functionTolerance = 1e-6f; p->gradientTolerance = 1e-8f;
  p->relaxTolerancesForCoarseLevels = 1; p->gradientEstimation = BPVO_B200_CD3; p->interp = BPVO_B200_LINEAR;
  p->lossFunction = BPVO_B200_TUKEY; p->descriptor = BPVO_B200_INTENSITY; p->verbosity = 0x20;
  p->minTranslationMagToKeyFrame = 0.15f; p->minRotationMagToKeyFrame = 5.0f;
  p->maxFractionOfGoodPointsToKeyFrame = 0.6f; p->goodPointThreshold = 0.85f;
  p->minNumPixelsForNonMaximaSuppression = 320 * 240; p->nonMaxSuppRadius = 1; p->minNumPixelsToWork = 256;
  p->minSaliency = 0.1f; p->minValidDisparity = 0.001f; p->maxValidDisparity = 512.0f;
  p->maxTestLevel = 0; p->withNormalization = 1;
  p->device_id = 0; p->flags = 0;
  p->dfSigma1 = 0.75f; p->dfSigma2 = 1.75f;
}

int bpvo_b200_device_count(void) {
  int n = 0;
  const cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); bp_fail(BPVO_B200_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e)); return 0; }
  return n;
}

void* bpvo_b200_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void bpvo_b200_host_free(void* p) { if (p) cudaFreeHost(p); }

// -------------------------------------------------------------------------------------------------
// ctx
// -------------------------------------------------------------------------------------------------
int bpvo_b200_create(bpvo_b200_ctx** out, const float K[9], float baseline, int rows, int cols, const bpvo_b200_params* p) {
  if (!out || !K || !p) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  if (p->numPyramidLevels <= 0 || p->numPyramidLevels > kMaxLevels)
    return bp_fail(BPVO_B200_ERR_INVALID_ARG, "invalid number of pyramid levels");          // dense_descriptor_pyramid.cc:37
  if (p->maxTestLevel < 0 || p->maxTestLevel >= p->numPyramidLevels)
    return bp_fail(BPVO_B200_ERR_INVALID_ARG, "invalid maxTestLevel");                       // dense_descriptor_pyramid.cc:38
  if (rows < 16 || cols < 24) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "image too small");
  if (p->descriptor != BPVO_B200_INTENSITY && p->descriptor != BPVO_B200_BITPLANES && p->descriptor != BPVO_B200_INTENSITY_AND_GRADIENT &&
      p->descriptor != BPVO_B200_DESCRIPTOR_FIELDS)
    return bp_fail(BPVO_B200_ERR_UNSUPPORTED, "DescriptorType 0x%x is not on the accelerated path (Intensity, IntensityAndGradient, DescriptorFieldsFirstOrder, BitPlanes)", p->descriptor);
  if (p->descriptor == BPVO_B200_INTENSITY_AND_GRADIENT && p->sigmaPriorToCensusTransform > 4.0f)
    return bp_fail(BPVO_B200_ERR_UNSUPPORTED, "GradientDescriptor: sigma > 4 needs a Gaussian of more than 33 taps");
  if (p->descriptor == BPVO_B200_DESCRIPTOR_FIELDS && (p->dfSigma1 > 16.0f || p->dfSigma2 > 16.0f))
    return bp_fail(BPVO_B200_ERR_UNSUPPORTED, "DescriptorFields: sigma > 16 needs a Gaussian of more than 33 taps");
  if (p->interp != BPVO_B200_LINEAR && p->interp != BPVO_B200_COSINE && p->interp != BPVO_B200_CUBIC && p->interp != BPVO_B200_CUBIC_HERMITE)
    return bp_fail(BPVO_B200_ERR_INVALID_ARG, "unknown InterpolationType");
  if (p->lossFunction != BPVO_B200_HUBER && p->lossFunction != BPVO_B200_TUKEY && p->lossFunction != BPVO_B200_L2)
    return bp_fail(BPVO_B200_ERR_INVALID_ARG, "unknown RobustFunction");
  if (p->gradientEstimation != BPVO_B200_CD3 && p->gradientEstimation != BPVO_B200_CD5)
    return bp_fail(BPVO_B200_ERR_INVALID_ARG, "unknown GradientEstimationType");
  // K is column-major: K[0]=fx K[4]=fy K[6]=cx K[7]=cy K[8]=1; skew / lower entries must be 0
  if (K[1] != 0.0f || K[2] != 0.0f || K[3] != 0.0f || K[5] != 0.0f || K[8] != 1.0f)
    return bp_fail(BPVO_B200_ERR_UNSUPPORTED, "K must be [fx 0 cx; 0 fy cy; 0 0 1]");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return bp_fail(BPVO_B200_ERR_CUDA, "no CUDA device: bpvo_b200 has no CPU fallback");
  }
  if (p->device_id < 0 || p->device_id >= ndev) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "device_id %d out of range (%d devices)", p->device_id, ndev);
  CUDA_TRY(cudaSetDevice(p->device_id));

  bpvo_b200_ctx* c = new bpvo_b200_ctx();
  const int rc = ctx_init(c, K, baseline, rows, cols, p);
  if (rc != BPVO_B200_OK) {                    // release whatever was allocated before the failure (the message survives)
    const std::string msg = g_err;
    bpvo_b200_destroy(c);
    g_err = msg;
    return rc;
  }
  *out = c;
  return BPVO_B200_OK;
}

}  // extern "C"

static int ctx_init(bpvo_b200_ctx* c, const float K[9], float baseline, int rows, int cols, const bpvo_b200_params* p) {
  c->p = *p; c->rows = rows; c->cols = cols;
  c->L = p->numPyramidLevels; c->baseline = baseline;
  if (getenv("BPVO_B200_NO_GRAPHS")) c->p.flags |= BPVO_B200_FLAG_NO_GRAPHS;      // A/B switches for measurements
  if (getenv("BPVO_B200_FAST_BLEND")) c->p.flags |= BPVO_B200_FLAG_FAST_BLEND;
  if (const char* e = getenv("BPVO_B200_SOLVER_CTAS")) c->solver_ctas = atoi(e);
  // test hook: start the exchange sequence numbers close to their wrap-around so that the reset paths get exercised
  if (const char* e = getenv("BPVO_B200_SEQ_INIT")) { c->ll_seq = (unsigned) strtoul(e, nullptr, 0); c->x_seq_init = c->ll_seq; }
  c->C = (p->descriptor == BPVO_B200_BITPLANES) ? 8 : (p->descriptor == BPVO_B200_DESCRIPTOR_FIELDS) ? 5 : (p->descriptor == BPVO_B200_INTENSITY_AND_GRADIENT) ? 3 : 1;
  c->CS = channel_stride(c->C);
  memcpy(c->K, K, sizeof(c->K));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, p->device_id));
  c->sm_count = prop.multiProcessorCount;
  c->coop = prop.cooperativeLaunch != 0;
  c->smem_optin = (int) prop.sharedMemPerBlockOptin;
  // per-level geometry: K halves (K(2,2) = 1), baseline doubles (vo_frame.cc:24-28); pyrDown sizes
  float fx = K[0], fy = K[4], cx = K[6], cy = K[7], b = baseline; int r = rows, cl = cols;
  for (int l = 0; l < c->L; ++l) {
    if (l > 0) { fx *= 0.5f; fy *= 0.5f; cx *= 0.5f; cy *= 0.5f; b *= 2.0f; r = (r + 1) / 2; cl = (cl + 1) / 2; }
    LevelGeom& g = c->geom[l];
    g.rows = r; g.cols = cl; g.fx = fx; g.fy = fy; g.cx = cx; g.cy = cy; g.Bf = b * fx;
    const int border = std::max(p->nonMaxSuppRadius, 3);
    g.capacity = std::max(0, r - 2 * border - 1) * std::max(0, cl - 2 * border - 1);
    g.capacity = (g.capacity + 15) & ~15;
    if (g.capacity == 0) g.capacity = 16;
  }
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreate(&c->ev0)); CUDA_TRY(cudaEventCreate(&c->ev1));
  CUDA_TRY(cudaEventCreate(&c->tm0)); CUDA_TRY(cudaEventCreate(&c->tm1));
  CUDA_TRY(cudaEventCreateWithFlags(&c->stage_free, cudaEventDisableTiming));
  CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&c->disp_done, cudaEventDisableTiming)); CUDA_TRY(cudaEventCreateWithFlags(&c->order_ev, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(c->disp_done, c->copy_stream));
  const size_t cap0 = (size_t) c->geom[c->p.maxTestLevel].capacity;
  size_t capmax = 0; for (int l = c->p.maxTestLevel; l < c->L; ++l) capmax = std::max(capmax, (size_t) c->geom[l].capacity);
  (void) cap0;
  CUDA_TRY(cudaMalloc(&c->work.res, capmax * c->CS * sizeof(float)));
  CUDA_TRY(cudaMalloc(&c->work.valid, capmax));
  CUDA_TRY(cudaMalloc(&c->work.hist, (kHistSets * kHistWords + 8 + 128) * sizeof(unsigned)));
  CUDA_TRY(cudaMalloc(&c->work.ll, (size_t) 4 * kMaxGrid * 32 * sizeof(uint4)));
  CUDA_TRY(cudaMemsetAsync(c->work.ll, 0, (size_t) 4 * kMaxGrid * 32 * sizeof(uint4), c->stream));
  CUDA_TRY(cudaMalloc(&c->work.msg, (size_t) 2 * kMaxGrid * kMsgWords * sizeof(uint4)));
  CUDA_TRY(cudaMemsetAsync(c->work.msg, 0, (size_t) 2 * kMaxGrid * kMsgWords * sizeof(uint4), c->stream));
  CUDA_TRY(cudaMalloc(&c->work.partials, (size_t) 1024 * kPartialStride * sizeof(double)));
  CUDA_TRY(cudaMalloc(&c->work.scale, sizeof(ScaleState)));
  CUDA_TRY(cudaMalloc(&c->d_mail, sizeof(Mailbox)));
  CUDA_TRY(cudaMemsetAsync(c->d_mail, 0, sizeof(Mailbox), c->stream));
  c->work.out = &c->d_mail->lin;
  CUDA_TRY(cudaMalloc(&c->work.ticket, 4 * sizeof(unsigned)));
  CUDA_TRY(cudaMalloc(&c->work.cand, ((size_t) kMaxGrid * kCandPerCta + kOvfCap + kSelList) * sizeof(float)));
  CUDA_TRY(cudaMalloc(&c->sel, sizeof(Sel)));
  CUDA_TRY(cudaMalloc(&c->export_buf, capmax * std::max(c->C * 7, 8) * sizeof(float)));      // Jacobian export / 32-byte point records
  CUDA_TRY(cudaMemsetAsync(c->work.hist, 0, (kHistSets * kHistWords + 8 + 128) * sizeof(unsigned), c->stream));
  CUDA_TRY(cudaMemsetAsync(c->work.ticket, 0, 4 * sizeof(unsigned), c->stream));
  CUDA_TRY(cudaMemsetAsync(c->sel, 0, sizeof(Sel), c->stream));
  k_reset_scale<<<1, 1, 0, c->stream>>>(c->work.scale);
  // template-build scratch (sized for level maxTestLevel = the largest one built)
  const int r0 = c->geom[c->p.maxTestLevel].rows, c0 = c->geom[c->p.maxTestLevel].cols;
  CUDA_TRY(cudaMalloc(&c->flags, (size_t) r0 * c0));
  if (c->C == 8 && p->sigmaPriorToCensusTransform > 0.0f) CUDA_TRY(cudaMalloc(&c->blur_tmp, (size_t) r0 * u8_pitch(c0)));   // pre-census blur output
  if (c->C == 3 || c->C == 5)        // f32 scratch planes of the gradient-based descriptors
    for (int k = 0; k < 5; ++k) CUDA_TRY(cudaMalloc(&c->plane[k], (size_t) r0 * c0 * sizeof(float)));
  CUDA_TRY(cudaMalloc(&c->block_counts, (size_t) (ceil_div(r0 * c0, kSelPerBlock) + 1) * sizeof(int)));
  CUDA_TRY(cudaMalloc(&c->hpartials, 1024 * 4 * sizeof(double)));
  CUDA_TRY(cudaMalloc(&c->hsums, 4 * sizeof(double)));
  // device outputs of estimate_pose + pinned mailboxes
  if (const char* e = getenv("BPVO_B200_TIMEOUT_MS")) c->timeout_ns = (unsigned long long) strtoull(e, nullptr, 0) * 1000000ull;
  CUDA_TRY(cudaMalloc(&c->d_prof, (64 + kMaxLevels * 16) * sizeof(long long)));      // [64] whole launch, then [level][16]
  CUDA_TRY(cudaMemsetAsync(c->d_prof, 0, (64 + kMaxLevels * 16) * sizeof(long long), c->stream));
  CUDA_TRY(cudaHostAlloc(&c->h_mail, sizeof(Mailbox), cudaHostAllocDefault));
  memset(c->h_mail, 0, sizeof(Mailbox));
  CUDA_TRY(cudaHostAlloc(&c->stage_img, (size_t) rows * cols, cudaHostAllocDefault));
  CUDA_TRY(cudaHostAlloc(&c->stage_disp, (size_t) rows * cols * sizeof(float), cudaHostAllocDefault));
  CUDA_TRY(bp_sync_stream(c));
  return BPVO_B200_OK;
}

extern "C" {

int bpvo_b200_destroy(bpvo_b200_ctx* c) {
  if (!c) return BPVO_B200_OK;
  cudaSetDevice(c->p.device_id);
  if (c->stream) bp_sync_stream(c);
  bp_comm_destroy(c);
  cudaFree(c->work.res); cudaFree(c->work.valid); cudaFree(c->work.hist); cudaFree(c->work.ll); cudaFree(c->work.msg); cudaFree(c->work.partials);
  cudaFree(c->work.scale); cudaFree(c->d_mail); cudaFree(c->work.ticket); cudaFree(c->work.cand); cudaFree(c->sel); cudaFree(c->export_buf);
  for (int k = 0; k < 5; ++k) cudaFree(c->plane[k]);
  cudaFree(c->flags); cudaFree(c->blur_tmp); cudaFree(c->block_counts); cudaFree(c->hpartials); cudaFree(c->hsums);
  cudaFree(c->d_prof); cudaFree(c->d_trace); cudaFree(c->d_trace_rows);
  if (c->h_mail) cudaFreeHost(c->h_mail); if (c->stage_img) cudaFreeHost(c->stage_img); if (c->stage_disp) cudaFreeHost(c->stage_disp);
  if (c->flush_buf) cudaFree(c->flush_buf);
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  for (cudaEvent_t e : {c->ev0, c->ev1, c->tm0, c->tm1, c->stage_free, c->disp_done, c->order_ev}) if (e) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  cudaGetLastError();
  delete c;
  return BPVO_B200_OK;
}

int bpvo_b200_synchronize(bpvo_b200_ctx* c) {
  if (!c) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null ctx");
  CUDA_TRY(bp_sync_stream(c));
  return BPVO_B200_OK;
}
int bpvo_b200_timer_start(bpvo_b200_ctx* c) {
  if (!c) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null ctx");
  CUDA_TRY(cudaEventRecord(c->tm0, c->stream));
  return BPVO_B200_OK;
}
int bpvo_b200_timer_stop(bpvo_b200_ctx* c, float* ms) {
  if (!c || !ms) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaEventRecord(c->tm1, c->stream));
  CUDA_TRY(cudaEventSynchronize(c->tm1));
  CUDA_TRY(cudaEventElapsedTime(ms, c->tm0, c->tm1));
  return BPVO_B200_OK;
}
int bpvo_b200_last_level_evals(bpvo_b200_ctx* c, int* evals) {
  if (!c || !evals) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  for (int l = 0; l < c->L; ++l) evals[l] = c->level_evals[l];
  return BPVO_B200_OK;
}
int bpvo_b200_get_phase_cycles(bpvo_b200_ctx* c, long long cycles[64], int reset) {
  if (!c || !cycles) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(c->p.device_id));
  CUDA_TRY(bp_sync_stream(c));
  CUDA_TRY(cudaMemcpy(cycles, c->d_prof, 64 * sizeof(long long), cudaMemcpyDeviceToHost));
  if (reset) CUDA_TRY(cudaMemset(c->d_prof, 0, 64 * sizeof(long long)));
  return BPVO_B200_OK;
}
int bpvo_b200_get_level_phase_cycles(bpvo_b200_ctx* c, long long* cycles /* [BPVO_B200_MAX_LEVELS][16] */, int reset) {
  if (!c || !cycles) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(c->p.device_id));
  CUDA_TRY(bp_sync_stream(c));
  CUDA_TRY(cudaMemcpy(cycles, c->d_prof + 64, kMaxLevels * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  if (reset) CUDA_TRY(cudaMemset(c->d_prof + 64, 0, kMaxLevels * 16 * sizeof(long long)));
  return BPVO_B200_OK;
}
int bpvo_b200_last_level_us(bpvo_b200_ctx* c, float* us) {
  if (!c || !us) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  for (int l = 0; l < c->L; ++l) us[l] = c->level_us[l];
  return BPVO_B200_OK;
}
// Throughput mode: the on-device GN loop of this ctx occupies only `ctas` SMs (one 256-thread CTA each; 0 = all SMs), so that
// several independent ctxs -- one VisualOdometry stream each, driven from different host threads -- run their solves side by side
// on one GPU (BASELINE.json configs[4] with more than one stream per GPU).  Results do not depend on the grid size beyond fp32
// summation order (tests/test_gpu_device_loop.py runs the parity checks at 1 ... 148 CTAs).
int bpvo_b200_set_solver_ctas(bpvo_b200_ctx* c, int ctas) {
  if (!c || ctas < 0) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "bad argument");
  c->solver_ctas = ctas;
  return BPVO_B200_OK;
}
int bpvo_b200_set_profiling(bpvo_b200_ctx* c, int enable) { if (!c) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null ctx"); c->profiling = enable != 0; return BPVO_B200_OK; }
int bpvo_b200_get_counters(bpvo_b200_ctx* c, bpvo_b200_counters* out) { if (!c || !out) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null"); *out = c->counters; return BPVO_B200_OK; }
int bpvo_b200_reset_counters(bpvo_b200_ctx* c) { if (!c) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null ctx"); memset(&c->counters, 0, sizeof(c->counters)); return BPVO_B200_OK; }

// -------------------------------------------------------------------------------------------------
// frame
// -------------------------------------------------------------------------------------------------
int bpvo_b200_frame_create(bpvo_b200_ctx* c, bpvo_b200_frame** out) {
  if (!c || !out) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(c->p.device_id));
  *out = nullptr;
  bpvo_b200_frame* f = new bpvo_b200_frame();
  f->ctx = c;
  const int rc = frame_init(c, f);
  if (rc != BPVO_B200_OK) {
    const std::string msg = g_err;
    bpvo_b200_frame_destroy(f);
    g_err = msg;
    return rc;
  }
  *out = f;
  return BPVO_B200_OK;
}

}  // extern "C"

static int frame_init(bpvo_b200_ctx* c, bpvo_b200_frame* f) {
  CUDA_TRY(cudaMalloc(&f->disp, (size_t) c->rows * c->cols * sizeof(float)));
  for (int l = 0; l < c->L; ++l) {
    const LevelGeom& g = c->geom[l];
    CUDA_TRY(cudaMalloc(&f->pyr[l], (size_t) g.rows * u8_pitch(g.cols)));
    CUDA_TRY(cudaMemsetAsync(f->pyr[l], 0, (size_t) g.rows * u8_pitch(g.cols), c->stream));
    if (l < c->p.maxTestLevel) continue;
    const size_t npx = (size_t) g.rows * g.cols, cap = (size_t) g.capacity;
    // one extra zero row (+ pad): the reference's cubic / Hermite footprint reaches row `rows` for yi = rows - 2
    // (photo_error.cc:358 bounds y by rows - 1 only) and reads past its buffer there; the engine reads zeros instead
    const size_t desc_elems = (npx + (size_t) g.cols + 4) * c->CS;
    CUDA_TRY(cudaMalloc(&f->desc[l], desc_elems * sizeof(float)));
    CUDA_TRY(cudaMemsetAsync(f->desc[l], 0, desc_elems * sizeof(float), c->stream));
    CUDA_TRY(cudaMalloc(&f->saliency[l], npx * sizeof(float)));
    CUDA_TRY(cudaMalloc(&f->pts[l], cap * sizeof(float4)));
    CUDA_TRY(cudaMalloc(&f->gx[l], cap * c->CS * sizeof(float)));
    CUDA_TRY(cudaMalloc(&f->gy[l], cap * c->CS * sizeof(float)));
    CUDA_TRY(cudaMalloc(&f->i0[l], cap * c->CS * sizeof(float)));
    CUDA_TRY(cudaMalloc(&f->inds[l], cap * sizeof(int)));
  }
  CUDA_TRY(cudaMalloc(&f->d_meta, kMaxLevels * sizeof(TemplateMeta)));
  CUDA_TRY(cudaMemsetAsync(f->d_meta, 0, kMaxLevels * sizeof(TemplateMeta), c->stream));
  CUDA_TRY(cudaHostAlloc(&f->h_meta, kMaxLevels * sizeof(TemplateMeta), cudaHostAllocDefault));
  memset(f->h_meta, 0, kMaxLevels * sizeof(TemplateMeta));
  CUDA_TRY(cudaEventCreateWithFlags(&f->meta_ready, cudaEventDisableTiming));
  // TMA tile-in / tile-out variant of the bit-planes descriptor kernel: built, parity-tested, and 14 % SLOWER than the plain
  // kernel in the same-box A/B (profiles/README.md) -- opt-in (BPVO_B200_FLAG_TMA_DESCRIPTOR or BPVO_B200_TMA=1)
  f->tma_ok = (c->C == 8) && ((c->p.flags & BPVO_B200_FLAG_TMA_DESCRIPTOR) || getenv("BPVO_B200_TMA")) && encode_tensor_maps(c, f);
  return BPVO_B200_OK;
}

extern "C" {

int bpvo_b200_frame_destroy(bpvo_b200_frame* f) {
  if (!f) return BPVO_B200_OK;
  bpvo_b200_ctx* c = f->ctx;
  cudaSetDevice(c->p.device_id);
  bp_sync_stream(c);
  cudaFree(f->disp);
  for (int l = 0; l < c->L; ++l) {
    cudaFree(f->pyr[l]); cudaFree(f->desc[l]); cudaFree(f->saliency[l]); cudaFree(f->pts[l]);
    cudaFree(f->gx[l]); cudaFree(f->gy[l]); cudaFree(f->i0[l]); cudaFree(f->inds[l]);
  }
  cudaFree(f->d_maps);
  cudaFree(f->d_meta); if (f->h_meta) cudaFreeHost(f->h_meta); if (f->meta_ready) cudaEventDestroy(f->meta_ready);
  for (int k = 0; k < 2; ++k) if (f->graph_exec[k]) cudaGraphExecDestroy(f->graph_exec[k]);
  if (c->last_ref == f) c->last_ref = nullptr;
  delete f;
  return BPVO_B200_OK;
}

// true if the DMA engine can read p directly: page-locked host memory or device / managed memory
static bool is_dma_able(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static int run_as_graph(bpvo_b200_ctx* c, bpvo_b200_frame* f, int which, int (*enqueue)(bpvo_b200_ctx*, bpvo_b200_frame*));

// cv::getGaussianKernel(ksize, sigma, CV_32F): exp in double -> float taps, normalised by the double sum of the float taps
static BlurTaps gaussian_taps(int ksize, double sigma) {
  BlurTaps t; memset(&t, 0, sizeof(t));
  const int half = ksize / 2;
  t.half = half;
  float k[33]; double sum = 0;
  const double sx = sigma > 0 ? sigma : ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8, sc = -0.5 / (sx * sx);
  for (int i = 0; i < ksize; ++i) { const double x = i - half; k[i] = (float) exp(sc * x * x); sum += k[i]; }
  sum = 1.0 / sum;
  for (int j = 0; j <= half; ++j) t.k[j] = (float) (k[half + j] * sum);
  return t;
}
static int enqueue_blur(bpvo_b200_ctx* c, const float* src, int rows, int cols, const BlurTaps& t, float* scratch, float* dst, int dst_stride, int dst_off) {
  const int nb = ceil_div(rows * cols, 256);
  blur_row_kernel<<<nb, 256, 0, c->stream>>>(src, rows, cols, t, scratch); LAUNCH_CHECK(c);
  blur_col_kernel<<<nb, 256, 0, c->stream>>>(scratch, rows, cols, t, dst, dst_stride, dst_off); LAUNCH_CHECK(c);
  return BPVO_B200_OK;
}
// GradientDescriptor::compute / DescriptorFields::compute (gradient_descriptor.cc:42-64, 101-116) of one pyramid level
static int enqueue_gradient_descriptor(bpvo_b200_ctx* c, bpvo_b200_frame* f, int l) {
  const LevelGeom& g = c->geom[l];
  const int npx = g.rows * g.cols, nb = ceil_div(npx, 256), pitch = u8_pitch(g.cols);
  float* A = c->plane[0]; float* Is = c->plane[1]; float* P = c->plane[2]; float* N = c->plane[3]; float* scratch = c->plane[4];
  u8_to_f32_kernel<<<nb, 256, 0, c->stream>>>(f->pyr[l], g.rows, g.cols, pitch, A); LAUNCH_CHECK(c);
  int rc;
  if (c->C == 3) {
    const float sg = c->p.sigmaPriorToCensusTransform;         // dense_descriptor.cc:49
    const float* src = A;
    if (sg > 0.0f) {                                            // cv::GaussianBlur(.., cv::Size(), sigma): ksize = cvRound(sigma * 8 + 1) | 1
      const int ksize = ((int) lrint((double) sg * 8 + 1)) | 1;
      if ((rc = enqueue_blur(c, A, g.rows, g.cols, gaussian_taps(ksize, (double) sg), scratch, Is, 1, 0))) return rc;
      src = Is;
    }
    gradient_descriptor_kernel<<<nb, 256, 0, c->stream>>>(f->pyr[l], g.rows, g.cols, pitch, src, f->desc[l]); LAUNCH_CHECK(c);
    return BPVO_B200_OK;
  }
  auto smooth_taps = [](float sigma) { const int k = std::max(5, 2 * (int) round((double) sigma) + 1); return gaussian_taps(k, (double) sigma); };   // imsmooth
  const float* src = A;
  if (c->p.dfSigma1 > 0.0f) { if ((rc = enqueue_blur(c, A, g.rows, g.cols, smooth_taps(c->p.dfSigma1), scratch, Is, 1, 0))) return rc; src = Is; }
  dfields_base_kernel<<<nb, 256, 0, c->stream>>>(f->pyr[l], g.rows, g.cols, pitch, f->desc[l]); LAUNCH_CHECK(c);
  for (int dir = 0; dir < 2; ++dir) {
    split_gradient_kernel<<<nb, 256, 0, c->stream>>>(src, g.rows, g.cols, dir, P, N); LAUNCH_CHECK(c);
    if (c->p.dfSigma2 > 0.0f) {
      const BlurTaps t = smooth_taps(c->p.dfSigma2);
      if ((rc = enqueue_blur(c, P, g.rows, g.cols, t, scratch, f->desc[l], 8, 1 + 2 * dir))) return rc;
      if ((rc = enqueue_blur(c, N, g.rows, g.cols, t, scratch, f->desc[l], 8, 2 + 2 * dir))) return rc;
    } else {
      plane_to_channel_kernel<<<nb, 256, 0, c->stream>>>(P, npx, f->desc[l], 8, 1 + 2 * dir); LAUNCH_CHECK(c);
      plane_to_channel_kernel<<<nb, 256, 0, c->stream>>>(N, npx, f->desc[l], 8, 2 + 2 * dir); LAUNCH_CHECK(c);
    }
  }
  return BPVO_B200_OK;
}

// the kernel sequence of setData: DenseDescriptorPyramid::init (dense_descriptor_pyramid.cc:67-78)
static int enqueue_descriptors(bpvo_b200_ctx* c, bpvo_b200_frame* f) {
  {
    PhaseTimer t(c, &c->counters.ms_pyramid);
    for (int l = 1; l < c->L; ++l) {
      const LevelGeom& s = c->geom[l - 1]; const LevelGeom& d = c->geom[l];
      pyr_down_kernel<<<dim3(ceil_div(d.cols, 32), ceil_div(d.rows, 8)), 256, 0, c->stream>>>(f->pyr[l - 1], s.rows, s.cols, u8_pitch(s.cols), f->pyr[l], d.rows, d.cols, u8_pitch(d.cols));
      LAUNCH_CHECK(c);
    }
  }
  {
    PhaseTimer t(c, &c->counters.ms_descriptor);
    for (int l = c->L - 1; l >= c->p.maxTestLevel; --l) {
      const LevelGeom& g = c->geom[l];
      if (c->C == 1) {
        intensity_kernel<<<ceil_div(g.rows * g.cols, 256), 256, 0, c->stream>>>(f->pyr[l], g.rows, g.cols, u8_pitch(g.cols), f->desc[l]);
      } else if (c->C == 3 || c->C == 5) {
        int rc = enqueue_gradient_descriptor(c, f, l);
        if (rc) return rc;
        continue;
      } else {
        float k[5]; double sum = 0; const double sg = c->p.sigmaBitPlanes > 0 ? (double) c->p.sigmaBitPlanes : 1.1;
        // cv::getGaussianKernel(5, sigma, CV_32F): exp in double -> float taps, normalised by the double sum of the float taps
        for (int i = 0; i < 5; ++i) { const double x = i - 2.0; k[i] = (float) exp(-0.5 / (sg * sg) * x * x); sum += k[i]; }
        for (int i = 0; i < 5; ++i) k[i] = (float) (k[i] * (1.0 / sum));
        const uint8_t* census_in = f->pyr[l];
        if (c->p.sigmaPriorToCensusTransform > 0.0f) {      // census.cc:63-65: cv::GaussianBlur(3x3) on the u8 level image first
          const double sc = (double) c->p.sigmaPriorToCensusTransform, e = exp(-0.5 / (sc * sc));
          const int ka = (int) lrint(256.0 * (e / (1.0 + 2.0 * e))), kc = 256 - 2 * ka;
          blur3_u8_kernel<<<dim3(ceil_div(g.cols, 32), ceil_div(g.rows, 8)), 256, 0, c->stream>>>(f->pyr[l], g.rows, g.cols, u8_pitch(g.cols), ka, kc, c->blur_tmp);
          LAUNCH_CHECK(c);
          census_in = c->blur_tmp;
        }
        if (f->tma_ok && census_in == f->pyr[l])       // tile in / out on the TMA engine (the tensor maps describe pyr[l] and desc[l])
          bitplanes_tma_kernel<<<dim3(ceil_div(g.cols, kTmTW), ceil_div(g.rows, kTmTH)), 256, 0, c->stream>>>(
              f->d_maps + 2 * l, f->d_maps + 2 * l + 1, g.rows, g.cols, k[2], k[3], k[4], c->p.sigmaBitPlanes > 0.0f ? 1 : 0);
        else
          bitplanes_kernel<<<dim3(ceil_div(g.cols, kBpTW), ceil_div(g.rows, kBpTH)), 256, 0, c->stream>>>(
              census_in, g.rows, g.cols, u8_pitch(g.cols), k[2], k[3], k[4], c->p.sigmaBitPlanes > 0.0f ? 1 : 0, f->desc[l]);
      }
      LAUNCH_CHECK(c);
    }
  }
  return BPVO_B200_OK;
}

// setData (vo_frame.cc:48-55) -> DenseDescriptorPyramid::init (dense_descriptor_pyramid.cc:67-78)
int bpvo_b200_frame_set_data(bpvo_b200_frame* f, const uint8_t* image, const float* disparity) {
  if (!f) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null frame");
  if (!image || !disparity) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "nullptr image/disparity");     // vo.cc:68
  bpvo_b200_ctx* c = f->ctx;
  CUDA_TRY(cudaSetDevice(c->p.device_id));
  const size_t npx = (size_t) c->rows * c->cols;
  {
    PhaseTimer t(c, &c->counters.ms_upload);
    const uint8_t* src_i = image; const float* src_d = disparity;
    const bool stage_i = !is_dma_able(image), stage_d = !is_dma_able(disparity);      // each on its own: the disparity map may live on the device (upstream stereo) while the image is pageable
    if (stage_i || stage_d) {
      CUDA_TRY(cudaEventSynchronize(c->stage_free));
      CUDA_TRY(cudaEventSynchronize(c->disp_done));
      if (stage_i) { memcpy(c->stage_img, image, npx); src_i = c->stage_img; }
      if (stage_d) { memcpy(c->stage_disp, disparity, npx * sizeof(float)); src_d = c->stage_disp; }
    }
    CUDA_TRY(cudaMemcpy2DAsync(f->pyr[0], (size_t) u8_pitch(c->cols), src_i, (size_t) c->cols, (size_t) c->cols, (size_t) c->rows, cudaMemcpyDefault, c->stream));
    CUDA_TRY(cudaEventRecord(c->stage_free, c->stream));
    // the disparity map: on the copy stream, behind everything `stream` holds so far (an earlier template build of this frame
    // object may still read the buffer), beside everything enqueued from here on
    CUDA_TRY(cudaEventRecord(c->order_ev, c->stream));
    CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->order_ev, 0));
    CUDA_TRY(cudaMemcpyAsync(f->disp, src_d, npx * sizeof(float), cudaMemcpyDefault, c->copy_stream));
    CUDA_TRY(cudaEventRecord(c->disp_done, c->copy_stream));
    c->disp_pending = true;
    c->counters.h2d_bytes += (int64_t) (npx * 5);
  }
  int rc = run_as_graph(c, f, 0, enqueue_descriptors);
  if (rc) return rc;
  f->has_data = true;
  return BPVO_B200_OK;
}

static LevelTemplate make_level_template(const bpvo_b200_frame* f, int l) {
  const LevelGeom& g = f->ctx->geom[l];
  LevelTemplate t;
  t.pts = f->pts[l]; t.gx = f->gx[l]; t.gy = f->gy[l]; t.i0 = f->i0[l]; t.meta = f->d_meta + l;
  t.fx = g.fx; t.fy = g.fy; t.cx = g.cx; t.cy = g.cy;
  return t;
}

// the launch sequence of setTemplate: TemplateData::setData per level (template_data.cc:37-142) + the header read-back
static int enqueue_template(bpvo_b200_ctx* c, bpvo_b200_frame* f) {
  for (int l = c->L - 1; l >= c->p.maxTestLevel; --l) {
    const LevelGeom& g = c->geom[l];
    const int npx = g.rows * g.cols;
    BP_SWITCH_C(c->C, saliency_kernel<CC><<<ceil_div(npx, 256), 256, 0, c->stream>>>(f->desc[l], g.rows, g.cols, f->saliency[l]));
    LAUNCH_CHECK(c);
    SelectArgs sa;
    sa.S = f->saliency[l]; sa.D = f->disp; sa.rows = g.rows; sa.cols = g.cols; sa.Dcols = c->cols; sa.level = l;
    sa.nms_radius = (npx >= c->p.minNumPixelsForNonMaximaSuppression) ? c->p.nonMaxSuppRadius : -1;   // template_data.cc:43-49
    sa.border = std::max(c->p.nonMaxSuppRadius, 3);
    sa.min_saliency = c->p.minSaliency; sa.min_disp = c->p.minValidDisparity; sa.max_disp = c->p.maxValidDisparity;
    const int nb = ceil_div(npx, kSelPerBlock);
    select_flags_kernel<<<nb, kSelThreads, 0, c->stream>>>(sa, c->flags, c->block_counts);
    LAUNCH_CHECK(c);
    select_scan_kernel<<<1, 1024, 0, c->stream>>>(c->block_counts, nb, f->d_meta + l, c->shard_rank, c->shard_size,
                                                   c->peer_mode ? c->shard_min_points : 0);
    LAUNCH_CHECK(c);
    PointArgs pa; pa.rows = g.rows; pa.cols = g.cols; pa.Dcols = c->cols; pa.level = l;
    pa.fx = g.fx; pa.fy = g.fy; pa.cx = g.cx; pa.cy = g.cy; pa.Bf = g.Bf;
    select_scatter_kernel<<<nb, kSelThreads, 0, c->stream>>>(pa, f->disp, c->flags, c->block_counts, f->d_meta + l, f->inds[l], f->pts[l]);
    LAUNCH_CHECK(c);
    if (c->p.withNormalization) {
      int rc = bp_hartley(c, f, l);
      if (rc != BPVO_B200_OK) return rc;
    } else {
      set_identity_normalization_kernel<<<1, 1, 0, c->stream>>>(f->d_meta + l);
      LAUNCH_CHECK(c);
    }
    const int rb = std::max(1, std::min(ceil_div(g.capacity * c->CS, 256), c->sm_count * 8));
    const int cd5 = c->p.gradientEstimation == BPVO_B200_CD5;
    BP_SWITCH_C(c->C, template_records_kernel<CC><<<rb, 256, 0, c->stream>>>(f->desc[l], g.cols, f->inds[l], f->d_meta + l, g.fx, g.fy, cd5, f->gx[l], f->gy[l], f->i0[l]));
    LAUNCH_CHECK(c);
  }
  CUDA_TRY(cudaMemcpyAsync(f->h_meta, f->d_meta, kMaxLevels * sizeof(TemplateMeta), cudaMemcpyDeviceToHost, c->stream));
  return BPVO_B200_OK;
}

// Replays a frame's fixed launch sequence as ONE CUDA graph launch (the ~36 small dependent kernels of a template build
// are launch-bound).  Captured per frame object on first use -- a frame's buffers never move, handles only swap roles.
// Not used while profiling (cudaEvent pairs per phase) or in the sharded mode (NCCL calls in the sequence).
static int run_as_graph(bpvo_b200_ctx* c, bpvo_b200_frame* f, int which, int (*enqueue)(bpvo_b200_ctx*, bpvo_b200_frame*)) {
  const bool graphable = c->shard_size <= 1 && !c->profiling && !(c->p.flags & BPVO_B200_FLAG_NO_GRAPHS) && !f->graph_failed[which];
  if (!graphable) return enqueue(c, f);
  if (!f->graph_exec[which]) {
    const int64_t before = c->counters.launches;
    cudaGraph_t g = nullptr;
    bool ok = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    int rc = BPVO_B200_OK;
    if (ok) {
      rc = enqueue(c, f);
      ok = cudaStreamEndCapture(c->stream, &g) == cudaSuccess && rc == BPVO_B200_OK && g != nullptr;
    }
    if (ok) ok = cudaGraphInstantiate(&f->graph_exec[which], g, 0) == cudaSuccess;
    if (g) cudaGraphDestroy(g);
    f->graph_launches[which] = (int) (c->counters.launches - before);
    c->counters.launches = before;
    if (!ok) {                         // capture not possible here: fall back to plain launches for this frame
      cudaGetLastError();
      f->graph_exec[which] = nullptr; f->graph_failed[which] = true;
      return enqueue(c, f);
    }
  }
  CUDA_TRY(cudaGraphLaunch(f->graph_exec[which], c->stream));
  c->counters.launches += f->graph_launches[which];
  return BPVO_B200_OK;
}

// setTemplate (vo_frame.cc:61-93); fully asynchronous
int bpvo_b200_frame_set_template(bpvo_b200_frame* f) {
  if (!f) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null frame");
  if (!f->has_data) return bp_fail(BPVO_B200_ERR_NO_DATA, "no data in frame");                       // vo_frame.cc:63
  bpvo_b200_ctx* c = f->ctx;
  CUDA_TRY(cudaSetDevice(c->p.device_id));
  PhaseTimer t(c, &c->counters.ms_template);
  CUDA_TRY(bp_join_copies(c));                      // the selection reads the disparity map
  int rc = run_as_graph(c, f, 1, enqueue_template);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(f->meta_ready, c->stream));
  c->counters.d2h_bytes += kMaxLevels * sizeof(TemplateMeta);
  f->has_template = true;
  return BPVO_B200_OK;
}

int bpvo_b200_frame_has_template(const bpvo_b200_frame* f) { return f && f->has_template; }
int bpvo_b200_frame_empty(const bpvo_b200_frame* f) { return !f || !f->has_data; }
int bpvo_b200_frame_clear(bpvo_b200_frame* f) { if (!f) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null frame"); f->has_data = false; f->has_template = false; return BPVO_B200_OK; }
int bpvo_b200_frame_num_levels(const bpvo_b200_frame* f) { return f ? f->ctx->L : 0; }
int bpvo_b200_frame_level_size(const bpvo_b200_frame* f, int level, int* rows, int* cols) {
  if (!f || level < 0 || level >= f->ctx->L) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "bad level");
  *rows = f->ctx->geom[level].rows; *cols = f->ctx->geom[level].cols;
  return BPVO_B200_OK;
}

static int wait_meta(const bpvo_b200_frame* f) {
  CUDA_TRY(cudaEventSynchronize(f->meta_ready));
  return BPVO_B200_OK;
}
#define CHECK_LEVEL(f, level)                                                                                           \
  if (!(f) || (level) < (f)->ctx->p.maxTestLevel || (level) >= (f)->ctx->L) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "bad frame/level")

int bpvo_b200_frame_num_points(const bpvo_b200_frame* f, int level, int* n) {
  CHECK_LEVEL(f, level);
  if (!f->has_template) { *n = 0; return BPVO_B200_OK; }
  int rc = wait_meta(f); if (rc) return rc;
  *n = f->h_meta[level].n;
  return BPVO_B200_OK;
}
int bpvo_b200_frame_get_points(const bpvo_b200_frame* f, int level, float* xyzw) {
  CHECK_LEVEL(f, level);
  if (!f->has_template) return bp_fail(BPVO_B200_ERR_NO_DATA, "frame has no template");
  int rc = wait_meta(f); if (rc) return rc;
  bpvo_b200_ctx* c = f->ctx;
  CUDA_TRY(cudaSetDevice(c->p.device_id));
  CUDA_TRY(cudaMemcpyAsync(xyzw, f->pts[level], (size_t) f->h_meta[level].n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(bp_sync_stream(c));
  c->counters.d2h_bytes += (int64_t) f->h_meta[level].n * 16;
  return BPVO_B200_OK;
}
int bpvo_b200_frame_get_pyramid(const bpvo_b200_frame* f, int level, uint8_t* out) {
  if (!f || level < 0 || level >= f->ctx->L) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "bad frame/level");
  bpvo_b200_ctx* c = f->ctx; const LevelGeom& g = c->geom[level];
  CUDA_TRY(cudaSetDevice(c->p.device_id));
  CUDA_TRY(cudaMemcpy2DAsync(out, (size_t) g.cols, f->pyr[level], (size_t) u8_pitch(g.cols), (size_t) g.cols, (size_t) g.rows, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(bp_sync_stream(c));
  return BPVO_B200_OK;
}
int bpvo_b200_frame_get_descriptor(const bpvo_b200_frame* f, int level, float* planes, int* channels) {
  CHECK_LEVEL(f, level);
  bpvo_b200_ctx* c = f->ctx; const LevelGeom& g = c->geom[level];
  CUDA_TRY(cudaSetDevice(c->p.device_id));
  const int npx = g.rows * g.cols;
  float* tmp = nullptr;
  CUDA_TRY(cudaMalloc(&tmp, (size_t) npx * c->C * sizeof(float)));
  deinterleave_kernel<<<ceil_div(npx, 256), 256, 0, c->stream>>>(f->desc[level], npx, c->C, c->CS, tmp);
  cudaError_t e = cudaMemcpyAsync(planes, tmp, (size_t) npx * c->C * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = bp_sync_stream(c);
  cudaFree(tmp);
  if (e != cudaSuccess) return bp_fail(BPVO_B200_ERR_CUDA, "descriptor download failed: %s", cudaGetErrorString(e));
  if (channels) *channels = c->C;
  return BPVO_B200_OK;
}
int bpvo_b200_frame_get_saliency(const bpvo_b200_frame* f, int level, float* out) {
  CHECK_LEVEL(f, level);
  bpvo_b200_ctx* c = f->ctx; const LevelGeom& g = c->geom[level];
  CUDA_TRY(cudaSetDevice(c->p.device_id));
  CUDA_TRY(cudaMemcpyAsync(out, f->saliency[level], (size_t) g.rows * g.cols * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(bp_sync_stream(c));
  return BPVO_B200_OK;
}
static int export_template(const bpvo_b200_frame* f, int level, float* pixels_out, float* J_out) {
  CHECK_LEVEL(f, level);
  if (!f->has_template) return bp_fail(BPVO_B200_ERR_NO_DATA, "frame has no template");
  int rc = wait_meta(f); if (rc) return rc;
  bpvo_b200_ctx* c = f->ctx;
  CUDA_TRY(cudaSetDevice(c->p.device_id));
  const int n = f->h_meta[level].n;
  if (n == 0) return BPVO_B200_OK;
  float* tmp = nullptr;
  CUDA_TRY(cudaMalloc(&tmp, (size_t) n * c->C * 7 * sizeof(float)));
  float* dJ = tmp; float* dP = tmp + (size_t) n * c->C * 6;
  LevelTemplate t = make_level_template(f, level);
  BP_SWITCH_C(c->C, export_template_kernel<CC><<<ceil_div(n * CC, 256), 256, 0, c->stream>>>(t, dP, dJ));
  cudaError_t e = cudaSuccess;
  if (pixels_out) e = cudaMemcpyAsync(pixels_out, dP, (size_t) n * c->C * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess && J_out) e = cudaMemcpyAsync(J_out, dJ, (size_t) n * c->C * 6 * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = bp_sync_stream(c);
  cudaFree(tmp);
  if (e != cudaSuccess) return bp_fail(BPVO_B200_ERR_CUDA, "template export failed: %s", cudaGetErrorString(e));
  return BPVO_B200_OK;
}
int bpvo_b200_frame_get_pixels(const bpvo_b200_frame* f, int level, float* out) { return export_template(f, level, out, nullptr); }
int bpvo_b200_frame_get_jacobians(const bpvo_b200_frame* f, int level, float* out) { return export_template(f, level, nullptr, out); }
int bpvo_b200_frame_get_point_inds(const bpvo_b200_frame* f, int level, int32_t* out) {
  CHECK_LEVEL(f, level);
  if (!f->has_template) return bp_fail(BPVO_B200_ERR_NO_DATA, "frame has no template");
  int rc = wait_meta(f); if (rc) return rc;
  bpvo_b200_ctx* c = f->ctx;
  CUDA_TRY(cudaSetDevice(c->p.device_id));
  CUDA_TRY(cudaMemcpyAsync(out, f->inds[level], (size_t) f->h_meta[level].n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(bp_sync_stream(c));
  return BPVO_B200_OK;
}
int bpvo_b200_frame_get_normalization(const bpvo_b200_frame* f, int level, float Tn[16]) {
  CHECK_LEVEL(f, level);
  if (!f->has_template) return bp_fail(BPVO_B200_ERR_NO_DATA, "frame has no template");
  int rc = wait_meta(f); if (rc) return rc;
  const TemplateMeta& m = f->h_meta[level];
  M44 T = identity44();
  T(0, 0) = T(1, 1) = T(2, 2) = m.s; T(0, 3) = -m.s * m.c1; T(1, 3) = -m.s * m.c2; T(2, 3) = -m.s * m.c3;
  memcpy(Tn, T.m, sizeof(T.m));
  return BPVO_B200_OK;
}

}  // extern "C"

// Hartley normalisation of one level: two tiny grid reductions (last-CTA fold) each followed by a one-thread finish;
// in the point-sharded mode the phase totals are all-reduced across ranks in between
int bp_hartley(bpvo_b200_ctx* c, bpvo_b200_frame* f, int l) {
  const int nb = std::max(1, std::min(ceil_div(c->geom[l].capacity, 256 * 8), c->sm_count * 2));
  for (int phase = 0; phase < 2; ++phase) {
    hartley_sum_kernel<<<nb, 256, 0, c->stream>>>(f->pts[l], f->d_meta + l, c->hpartials, c->work.ticket + 1, c->hsums, phase, c->shard_rank);
    LAUNCH_CHECK(c);
    int rc = bp_comm_allreduce_f64(c, c->hsums, 3);
    if (rc) return rc;
    hartley_finish_kernel<<<1, 1, 0, c->stream>>>(f->d_meta + l, c->hsums, phase);
    LAUNCH_CHECK(c);
  }
  return BPVO_B200_OK;
}

// -------------------------------------------------------------------------------------------------
// linearize (fine seam) and estimate_pose (coarse seam)
// -------------------------------------------------------------------------------------------------
static int lin_grid(const bpvo_b200_ctx* c, const bpvo_b200_frame* ref, int level, int ctas_per_sm) {
  // host-driven kernels: as many 256-thread CTAs as their register use lets an SM hold (k_residuals / k_reduce: 2,
  // k_select: 4) when the level has the points to fill them (the template's size once its header has arrived on the
  // host, its upper bound before that); one CTA per SM otherwise
  int n = c->geom[level].capacity;
  if (ref && ref->has_template) {
    if (cudaEventQuery(ref->meta_ready) == cudaSuccess) n = ref->h_meta[level].n;
    else cudaGetLastError();           // cudaErrorNotReady must not be mistaken for a launch failure later
  }
  const int need = ceil_div(n, kLinThreads);
  const int per_sm = (need >= 2 * c->sm_count * ctas_per_sm) ? ctas_per_sm : 1;
  return std::max(1, std::min(std::min(c->sm_count * per_sm, 1024), need));
}

template <int C, int BLEND>
static int launch_linearize_t(bpvo_b200_ctx* c, const bpvo_b200_frame* ref, const bpvo_b200_frame* cur, int level, const M44& T) {
  LinArgs a;
  a.tmpl = make_level_template(ref, level);
  a.img.desc = cur->desc[level]; a.img.rows = c->geom[level].rows; a.img.cols = c->geom[level].cols;
  a.work = c->work; a.loss = c->p.lossFunction; a.interp = c->p.interp; a.good_thr = c->p.goodPointThreshold;
  a.hset = c->work.hist; a.sel = c->sel;
  make_projection(a.tmpl, T, a.P);
  const int grid = lin_grid(c, ref, level, 2), grid_sel = lin_grid(c, ref, level, 4);
  const bool robust = c->p.lossFunction != BPVO_B200_L2;
  if (robust) CUDA_TRY(cudaMemsetAsync(c->work.hist, 0, kHistWords * sizeof(unsigned), c->stream));
  k_residuals<C, BLEND><<<grid, kLinThreads, 0, c->stream>>>(a); LAUNCH_CHECK(c);
  // point-sharded mode: the histograms and the 30 sums are all-reduced -- unless this level is REPLICATED (peer-memory mode keeps
  // levels under shard_min_points whole on every rank): every rank then already holds all points and a sum over ranks would
  // count each of them nranks times
  bool sharded = c->shard_size > 1;
  int rc;
  if (sharded && c->peer_mode) {
    if ((rc = wait_meta(ref))) return rc;
    if (ref->h_meta[level].replicated) sharded = false;
  }
  if (robust) {
    if (sharded && (rc = bp_comm_allreduce_u32(c, a.hset, kHist1Bins))) return rc;                       // global level-1 histogram
    k_select<C, 2><<<grid_sel, kLinThreads, 0, c->stream>>>(a); LAUNCH_CHECK(c);
    if (sharded && (rc = bp_comm_allreduce_u32(c, a.hset + kHist1Bins, 2 * kHist2Bins))) return rc;
    k_select<C, 3><<<grid_sel, kLinThreads, 0, c->stream>>>(a); LAUNCH_CHECK(c);
    if (sharded && (rc = bp_comm_allreduce_u32(c, a.hset + kHist1Bins + 2 * kHist2Bins, 2 * kHist3Bins))) return rc;
  }
  if (!sharded) {
    k_reduce<C><<<grid, kLinThreads, 0, c->stream>>>(a); LAUNCH_CHECK(c);
  } else {
    k_reduce_sharded<C><<<grid, kLinThreads, 0, c->stream>>>(a, c->comm_buf); LAUNCH_CHECK(c);
    if ((rc = bp_comm_allreduce_f64(c, c->comm_buf, 32))) return rc;                                     // the 30 normal-equation sums
    k_finalize_sums<<<1, 32, 0, c->stream>>>(c->comm_buf, c->work.out); LAUNCH_CHECK(c);
  }
  c->counters.linearize_calls++;
  c->last_ref = ref; c->last_level = level;
  return BPVO_B200_OK;
}

// arithmetic of the bilinear blend (phase_residuals): the reference's double expression unless the ctx asks for fp32 FMAs on
// bit-planes (BPVO_B200_FLAG_FAST_BLEND); intensity always computes in fp64
template <int C>
static int launch_linearize(bpvo_b200_ctx* c, const bpvo_b200_frame* ref, const bpvo_b200_frame* cur, int level, const M44& T) {
  if constexpr (C == 8) { if (c->p.flags & BPVO_B200_FLAG_FAST_BLEND) return launch_linearize_t<C, 1>(c, ref, cur, level, T); }
  return launch_linearize_t<C, 0>(c, ref, cur, level, T);
}

static int check_pair(bpvo_b200_ctx* c, const bpvo_b200_frame* ref, const bpvo_b200_frame* cur) {
  if (!c || !ref || !cur) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  if (ref->ctx != c || cur->ctx != c) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "frames belong to another ctx");
  if (!ref->has_template) return bp_fail(BPVO_B200_ERR_NO_POINTS, "you should call setData before calling computeResiduals");
  if (!cur->has_data) return bp_fail(BPVO_B200_ERR_NO_DATA, "no data in frame");
  return BPVO_B200_OK;
}

extern "C" int bpvo_b200_linearize(bpvo_b200_ctx* c, const bpvo_b200_frame* ref, const bpvo_b200_frame* cur, int level,
                                    const float T[16], int first_call_of_level, float H[36], float G[6], float* f_norm, float* sigma, int* n_valid) {
  int rc = check_pair(c, ref, cur); if (rc) return rc;
  if (level < c->p.maxTestLevel || level >= c->L || !T) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "bad level / pose");
  CUDA_TRY(cudaSetDevice(c->p.device_id));
  {
    PhaseTimer t(c, &c->counters.ms_linearize);
    if (first_call_of_level) { k_reset_scale<<<1, 1, 0, c->stream>>>(c->work.scale); LAUNCH_CHECK(c); }
    M44 Tm; memcpy(Tm.m, T, sizeof(Tm.m));
    BP_SWITCH_C(c->C, rc = launch_linearize<CC>(c, ref, cur, level, Tm));
    if (rc) return rc;
  }
  CUDA_TRY(cudaMemcpyAsync(&c->h_mail->lin, c->work.out, sizeof(LinOut), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(bp_sync_stream(c));
  c->counters.d2h_bytes += sizeof(LinOut);
  const LinOut& o = c->h_mail->lin;
  if (H) memcpy(H, o.H, sizeof(o.H));
  if (G) memcpy(G, o.G, sizeof(o.G));
  if (f_norm) *f_norm = o.f_norm;
  if (sigma) *sigma = o.sigma;
  if (n_valid) *n_valid = o.n_valid;
  if (ref->h_meta[level].n_total == 0 && cudaEventQuery(ref->meta_ready) == cudaSuccess)
    return bp_fail(BPVO_B200_ERR_NO_POINTS, "you should call setData before calling computeResiduals");   // template_data.cc:177
  return BPVO_B200_OK;
}

// PoseEstimatorBase::run driven from the host on top of the fine seam (pose_estimator_base.h:324-407)
static int host_run_level(bpvo_b200_ctx* c, const bpvo_b200_frame* ref, const bpvo_b200_frame* cur, int level, M44& T, bpvo_b200_stats& st, int& evals) {
  int rc = wait_meta(ref); if (rc) return rc;
  const TemplateMeta& m = ref->h_meta[level];
  if (m.n_total == 0) return bp_fail(BPVO_B200_ERR_NO_POINTS, "you should call setData before calling computeResiduals");
  M44 Tn = identity44();
  Tn(0, 0) = Tn(1, 1) = Tn(2, 2) = m.s; Tn(0, 3) = -m.s * m.c1; Tn(1, 3) = -m.s * m.c2; Tn(2, 3) = -m.s * m.c3;
  const M44 Tn_inv = inverse44(Tn);
  const float sqrt_eps = sqrtf(FLT_EPSILON);
  st.numIterations = 0; st.firstOrderOptimality = 0; st.status = BPVO_B200_MAX_ITERS; st.finalError = -1.0f;
  float H[36], G[6], dp[6], ndp[6], f_norm = 0, g_norm = 0;
  int n_evals = 0;
  auto gnorm = [&]() { float g = 0; for (int k = 0; k < 6; ++k) g = std::max(g, fabsf(G[k])); return g; };
  M44 Td = T;
  rc = bpvo_b200_linearize(c, ref, cur, level, Td.m, 1, H, G, &f_norm, nullptr, nullptr); if (rc) return rc;
  ++n_evals;
  g_norm = gnorm();
  const float g_tol = c->p.gradientTolerance * std::max(g_norm, sqrt_eps);
  if (g_norm < g_tol) { st.status = BPVO_B200_GRAD_TOL; st.finalError = f_norm; st.numIterations = 1; st.firstOrderOptimality = g_norm; evals += n_evals; c->level_evals[level] = n_evals; return BPVO_B200_OK; }
  if (!solve6(H, G, dp)) { st.status = BPVO_B200_SOLVER_ERROR; st.finalError = f_norm; evals += n_evals; c->level_evals[level] = n_evals; return BPVO_B200_OK; }
  float f_prev = 0.0f, dp_prev = 0.0f; bool conv = false;
  for (int k = 0; k < 6; ++k) ndp[k] = -dp[k];
  Td = mul44(Td, params_to_pose(Tn, Tn_inv, ndp));
  do {
    float dpn = 0; for (int k = 0; k < 6; ++k) dpn += dp[k] * dp[k]; dpn = sqrtf(dpn);
    g_norm = gnorm();
    if (dpn < c->p.parameterTolerance || dpn < c->p.parameterTolerance * (sqrt_eps + dp_prev)) { st.status = BPVO_B200_PARAM_TOL; conv = true; }
    else if (f_norm < c->p.functionTolerance || f_norm < c->p.functionTolerance * (sqrt_eps + f_prev) ||
             fabsf(f_norm - f_prev) < c->p.functionTolerance) { st.status = BPVO_B200_FUNC_TOL; conv = true; }
    else if (g_norm < g_tol) { st.status = BPVO_B200_GRAD_TOL; conv = true; }
    dp_prev = dpn; f_prev = f_norm;
    if (!conv) {
      rc = bpvo_b200_linearize(c, ref, cur, level, Td.m, 0, H, G, &f_norm, nullptr, nullptr); if (rc) return rc;
      ++n_evals;
      if (!solve6(H, G, dp)) { st.status = BPVO_B200_SOLVER_ERROR; break; }
    }
    for (int k = 0; k < 6; ++k) ndp[k] = -dp[k];
    Td = mul44(Td, params_to_pose(Tn, Tn_inv, ndp));       // also on the converged pass (Q1)
  } while (st.numIterations++ < c->p.maxIterations && !conv && n_evals < 1200);
  if (st.status != BPVO_B200_SOLVER_ERROR) T = Td;
  st.numIterations -= 1; st.finalError = f_norm; st.firstOrderOptimality = g_norm;
  evals += n_evals; c->level_evals[level] = n_evals;
  return BPVO_B200_OK;
}

// optional overrides of a persistent-kernel launch (parity hooks, bpvo_b200_debug_device_linearize)
struct SolveOverride {
  DebugArgs dbg{};
  int grid = 0;            // 0 = one CTA per SM
  int cache_bytes = -1;    // -1 = everything the SM has
};

template <int C, int BLEND, bool PEER, int FIX>
static int launch_estimate_pose_t(bpvo_b200_ctx* c, const bpvo_b200_frame* ref, const bpvo_b200_frame* cur, const M44& T_init, const SolveOverride* ov,
                                  int lvl_first = -1, int lvl_last = -1, int chain = 0) {
  SolveArgs a;
  memset(&a, 0, sizeof(a));
  a.lvl_first = lvl_first >= 0 ? lvl_first : c->L - 1; a.lvl_last = lvl_last >= 0 ? lvl_last : c->p.maxTestLevel; a.chain = chain;
  if (ov) a.dbg = ov->dbg;
  if (c->d_trace && !(ov && ov->dbg.n > 0)) { a.dbg.trace = c->d_trace; a.dbg.trace_cap = kTraceRows; a.dbg.trace_rows = c->d_trace_rows; }
  for (int l = c->p.maxTestLevel; l < c->L; ++l) {
    a.tmpl[l] = make_level_template(ref, l);
    a.img[l].desc = cur->desc[l]; a.img[l].rows = c->geom[l].rows; a.img[l].cols = c->geom[l].cols;
  }
  a.sp.max_iterations = c->p.maxIterations; a.sp.max_fun_evals = 1200;
  a.sp.parameter_tolerance = c->p.parameterTolerance; a.sp.function_tolerance = c->p.functionTolerance; a.sp.gradient_tolerance = c->p.gradientTolerance;
  a.sp.loss = c->p.lossFunction; a.sp.interp = c->p.interp; a.sp.good_threshold = c->p.goodPointThreshold;
  a.sp.max_test_level = c->p.maxTestLevel; a.sp.num_levels = c->L;
  a.work = c->work; a.T_init = T_init; a.T_out = &c->d_mail->T; a.stats = c->d_mail->stats; a.num_fun_evals = &c->d_mail->evals; a.aborted_out = &c->d_mail->aborted;
  a.timeout_ns = c->timeout_ns;
  a.prof = c->profiling ? c->d_prof : nullptr;
  a.prof_lvl = c->profiling ? c->d_prof + 64 : nullptr;
  Sel* sel = c->sel;
  // the grid-barrier counter (behind the histogram sets) must be zero on entry; the kernel zeroes the sets itself
  CUDA_TRY(cudaMemsetAsync(c->work.hist + (size_t) kHistSets * kHistWords, 0, 8 * sizeof(unsigned), c->stream));
  // sequence numbers of the flag-in-data exchanges: unique per exchange over the life of the mailboxes
  const unsigned seq_span = (unsigned) (c->L * (std::min(c->p.maxIterations + 2, 1200) + 2) + 2) + (unsigned) (ov ? ov->dbg.n : 0);
  if (c->ll_seq == 0 || c->ll_seq > 0xffffffffu - seq_span) {
    CUDA_TRY(cudaMemsetAsync(c->work.ll, 0, (size_t) 4 * kMaxGrid * 32 * sizeof(uint4), c->stream));
    CUDA_TRY(cudaMemsetAsync(c->work.msg, 0, (size_t) 2 * kMaxGrid * kMsgWords * sizeof(uint4), c->stream));
    c->ll_seq = 1;
  }
  a.seq_base = c->ll_seq; c->ll_seq += seq_span;
  if (c->peer_mode) {
    // cross-rank exchanges: at most 6 per linearize (histogram + list + sums, or 3 radix histograms + sums); every rank
    // advances identically (lock-step launches).  Wrap-around would need a collective reset of the mailboxes: refuse instead.
    const unsigned xspan = seq_span * 8u;
    if (c->x_seq > 0x7fffffffu) {
      // collective reset (all ranks reach this at the same launch): barrier, clear the own mailbox, barrier, restart at 1
      int rc = bp_comm_allreduce_f64(c, c->comm_buf, 1); if (rc) return rc;
      CUDA_TRY(cudaMemsetAsync(c->xbox, 0, (size_t) 2 * kXRanks * kXWords * sizeof(uint2), c->stream));
      rc = bp_comm_allreduce_f64(c, c->comm_buf, 1); if (rc) return rc;
      c->x_seq = 1;
    }
    a.peer.rank = c->shard_rank; a.peer.nranks = c->shard_size; a.peer.xseq_base = c->x_seq; a.peer.lbox = c->lbox;
    for (int r = 0; r < c->shard_size; ++r) a.peer.box[r] = c->xpeer[r];
    c->x_seq += xspan;
  }
  // dynamic shared memory: candidate scratch + everything else the SM has for the template cache (1 CTA per SM; the
  // kernel plans per level which fields fit)
  int cache_bytes = ((c->smem_optin - 24 * 1024 - kScratchBytes) / 1024) * 1024;
  cache_bytes = std::max(0, cache_bytes);
  if (ov && ov->cache_bytes >= 0) cache_bytes = std::min(cache_bytes, ov->cache_bytes);
  const size_t dyn = (size_t) kScratchBytes + (size_t) cache_bytes;
  {                                      // the attribute is per kernel instantiation and per device (a process may drive several of either)
    static std::atomic<size_t> configured[64];      // (two host threads may both make the call: harmless)
    const int dev = c->p.device_id & 63;
    if (configured[dev].load(std::memory_order_relaxed) != dyn) {
      CUDA_TRY(cudaFuncSetAttribute((const void*) k_estimate_pose<C, BLEND, PEER, FIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dyn));
      configured[dev].store(dyn, std::memory_order_relaxed);
    }
  }
  void* args[] = {&a, &sel, &cache_bytes};
  int grid = std::min(c->sm_count, kMaxGrid);
  if (c->solver_ctas > 0) grid = std::min(grid, c->solver_ctas);     // throughput mode: several ctxs share the SMs
  if (ov && ov->grid > 0) grid = std::min(grid, ov->grid);
  CUDA_TRY(cudaLaunchCooperativeKernel((void*) k_estimate_pose<C, BLEND, PEER, FIX>, dim3(grid), dim3(kLinThreads), args, dyn, c->stream));
  c->counters.launches++;
  return BPVO_B200_OK;
}
template <int C>
static int launch_estimate_pose(bpvo_b200_ctx* c, const bpvo_b200_frame* ref, const bpvo_b200_frame* cur, const M44& T_init, const SolveOverride* ov = nullptr) {
  // (the peer-memory instantiation carries the cross-rank exchanges; the single-GPU one is a third smaller)
  if constexpr (C == 8) {
    if (c->p.flags & BPVO_B200_FLAG_FAST_BLEND) return launch_estimate_pose_t<C, 1, false, 0>(c, ref, cur, T_init, ov);
    // the bit-planes workloads' hot configurations with the loss function and the linear interpolant compiled in
    if (!c->peer_mode && c->p.interp == BPVO_B200_LINEAR && !getenv("BPVO_B200_GENERIC_KERNEL")) {
      if (c->p.lossFunction == BPVO_B200_TUKEY) {
        // finest level too large for the shared-memory cache (1080p dense: 48 points per thread): the instantiation that keeps
        // two points per thread in flight in the residual phase.  Decided from the template header when it is already on the
        // host (never waits for it); results are bit-identical either way.
        bool streams = false;
        if (!getenv("BPVO_B200_NO_STREAM2") && cudaEventQuery(ref->meta_ready) == cudaSuccess) {
          int grid = std::min(c->sm_count, kMaxGrid);
          if (c->solver_ctas > 0) grid = std::min(grid, c->solver_ctas);
          if (ov && ov->grid > 0) grid = std::min(grid, ov->grid);
          const int n = ref->h_meta[c->p.maxTestLevel].n;
          int cache_bytes = std::max(0, ((c->smem_optin - 24 * 1024 - kScratchBytes) / 1024) * 1024);
          if (ov && ov->cache_bytes >= 0) cache_bytes = std::min(cache_bytes, ov->cache_bytes);      // (the parity hook can force streaming at small sizes)
          const TplCache plan = tpl_cache_plan<C>((unsigned) kScratchBytes, cache_bytes, (n + grid * kLinThreads - 1) / (grid * kLinThreads));
          streams = plan.K > 1 && plan.pts == kTcNone && plan.f[TC_R] == kTcNone;
        } else {
          cudaGetLastError();
        }
        if (streams) {
          // the solve split by level: the cached levels in the usual kernel, then the streaming level in its own instantiation,
          // which starts from the pose the first launch left on the device
          const int fine = c->p.maxTestLevel;
          if (ov && ov->dbg.n > 0)           // parity hook (one level): the kernel that level would run in
            return ov->dbg.level == fine ? launch_estimate_pose_t<C, 0, false, BPVO_B200_TUKEY | 0x100>(c, ref, cur, T_init, ov)
                                         : launch_estimate_pose_t<C, 0, false, BPVO_B200_TUKEY>(c, ref, cur, T_init, ov);
          if (c->L - 1 > fine) {
            const int rc = launch_estimate_pose_t<C, 0, false, BPVO_B200_TUKEY>(c, ref, cur, T_init, ov, c->L - 1, fine + 1, 0);
            if (rc != BPVO_B200_OK) return rc;
            return launch_estimate_pose_t<C, 0, false, BPVO_B200_TUKEY | 0x100>(c, ref, cur, T_init, ov, fine, fine, 1);
          }
          return launch_estimate_pose_t<C, 0, false, BPVO_B200_TUKEY | 0x100>(c, ref, cur, T_init, ov);
        }
        return launch_estimate_pose_t<C, 0, false, BPVO_B200_TUKEY>(c, ref, cur, T_init, ov);
      }
      if (c->p.lossFunction == BPVO_B200_HUBER) return launch_estimate_pose_t<C, 0, false, BPVO_B200_HUBER>(c, ref, cur, T_init, ov);
    }
  }
  if (c->peer_mode) return launch_estimate_pose_t<C, 0, true, 0>(c, ref, cur, T_init, ov);
  return launch_estimate_pose_t<C, 0, false, 0>(c, ref, cur, T_init, ov);
}

extern "C" int bpvo_b200_estimate_pose(bpvo_b200_ctx* c, const bpvo_b200_frame* ref, const bpvo_b200_frame* cur,
                                        const float T_init[16], float T_est[16], bpvo_b200_stats* stats, int* num_fun_evals) {
  int rc = check_pair(c, ref, cur); if (rc) return rc;
  if (!T_init || !T_est || !stats) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(c->p.device_id));
  for (int l = 0; l < c->L; ++l) { stats[l].numIterations = 0; stats[l].finalError = -1.0f; stats[l].firstOrderOptimality = -1.0f; stats[l].status = BPVO_B200_SOLVER_ERROR; }   // types.cc:306-310
  M44 T; memcpy(T.m, T_init, sizeof(T.m));
  int evals = 0;
  c->counters.solve_calls++;
  const bool host_loop = (c->p.flags & BPVO_B200_FLAG_HOST_SOLVE) || !c->coop || (c->shard_size > 1 && !c->peer_mode);
  if (host_loop) {
    for (int l = c->L - 1; l >= c->p.maxTestLevel; --l) {              // vo_pose_estimator.cc:76-84
      rc = host_run_level(c, ref, cur, l, T, stats[l], evals);
      if (rc) return rc;
    }
  } else {
    {
      PhaseTimer t(c, &c->counters.ms_linearize);
      BP_SWITCH_C(c->C, rc = launch_estimate_pose<CC>(c, ref, cur, T));
      if (rc) return rc;
    }
    Mailbox* mb = c->h_mail;
    CUDA_TRY(cudaMemcpyAsync(mb, c->d_mail, sizeof(Mailbox), cudaMemcpyDeviceToHost, c->stream));      // pose, statistics, LinOut: one D2H
    CUDA_TRY(bp_sync_stream(c));
    c->counters.d2h_bytes += sizeof(Mailbox);
    if (mb->aborted) return bp_fail(BPVO_B200_ERR_CUDA, "on-device GN loop: a grid barrier / exchange timed out");
    T = mb->T; evals = mb->evals;
    c->counters.linearize_calls += evals;
    for (int l = c->p.maxTestLevel; l < c->L; ++l) {
      if (mb->stats[l].status == -3) return bp_fail(BPVO_B200_ERR_NO_POINTS, "you should call setData before calling computeResiduals");
      if (mb->stats[l].status == -4) return bp_fail(BPVO_B200_ERR_CUDA, "on-device GN loop: a grid barrier / exchange timed out");
      stats[l].numIterations = mb->stats[l].num_iterations; stats[l].finalError = mb->stats[l].final_error;
      stats[l].firstOrderOptimality = mb->stats[l].first_order_optimality; stats[l].status = mb->stats[l].status;
      c->level_evals[l] = mb->stats[l].num_evals; c->level_us[l] = mb->stats[l].us;
    }
    c->last_ref = ref; c->last_level = c->p.maxTestLevel;
  }
  memcpy(T_est, T.m, sizeof(T.m));
  if (num_fun_evals) *num_fun_evals = evals;
  return BPVO_B200_OK;
}

// getWeights / residuals / valid of the last linearize, in the reference's channel-major layout.
// *count: in = capacity of the caller's buffer in floats (ignored when the buffer is NULL), out = total C*N.
// Only min(capacity, total) entries are computed-and-copied; the first N entries are channel 0 (what vo.cc:264 uses).
static int export_last(bpvo_b200_ctx* c, float* w, float* r, size_t* count) {
  if (!c || !count) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  const size_t capacity = *count;
  if (!c->last_ref) { *count = 0; return BPVO_B200_OK; }
  int rc = wait_meta(c->last_ref); if (rc) return rc;
  const int n = c->last_ref->h_meta[c->last_level].n;
  *count = (size_t) n * c->C;
  if ((!w && !r) || n == 0) return BPVO_B200_OK;
  const size_t ncopy = (capacity == 0 || capacity > *count) ? *count : capacity;
  CUDA_TRY(cudaSetDevice(c->p.device_id));
  // sigma of the last linearize
  CUDA_TRY(cudaMemcpyAsync(&c->h_mail->lin, c->work.out, sizeof(LinOut), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(bp_sync_stream(c));
  const float sigma = c->h_mail->lin.sigma;
  float* dw = w ? c->export_buf : nullptr;
  float* dr = r ? c->export_buf + (size_t) n * c->C : nullptr;
  BP_SWITCH_C(c->C, k_export_weights<CC><<<ceil_div(n * CC, 256), 256, 0, c->stream>>>(c->work.res, n, sigma, c->p.lossFunction, dw, dr));
  LAUNCH_CHECK(c);
  if (w) CUDA_TRY(cudaMemcpyAsync(w, dw, ncopy * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (r) CUDA_TRY(cudaMemcpyAsync(r, dr, ncopy * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(bp_sync_stream(c));
  c->counters.d2h_bytes += (int64_t) (ncopy * sizeof(float) * ((w ? 1 : 0) + (r ? 1 : 0)));
  return BPVO_B200_OK;
}
extern "C" int bpvo_b200_get_weights(bpvo_b200_ctx* c, float* w, size_t* count) { return export_last(c, w, nullptr, count); }
extern "C" int bpvo_b200_get_residuals(bpvo_b200_ctx* c, float* r, size_t* count) { return export_last(c, nullptr, r, count); }
extern "C" int bpvo_b200_get_valid(bpvo_b200_ctx* c, uint8_t* v, size_t* count) {
  if (!c || !count) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  if (!c->last_ref) { *count = 0; return BPVO_B200_OK; }
  int rc = wait_meta(c->last_ref); if (rc) return rc;
  *count = (size_t) c->last_ref->h_meta[c->last_level].n;
  if (!v || *count == 0) return BPVO_B200_OK;
  CUDA_TRY(cudaSetDevice(c->p.device_id));
  CUDA_TRY(cudaMemcpyAsync(v, c->work.valid, *count, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(bp_sync_stream(c));
  return BPVO_B200_OK;
}

// getFractionOfGoodPoints (vo_pose_estimator.cc:101-107): counted inside the reduce phase, no bulk D2H
extern "C" int bpvo_b200_fraction_good(bpvo_b200_ctx* c, float thresh, float* frac) {
  if (!c || !frac) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  if (!c->last_ref) { *frac = 0.0f; return BPVO_B200_OK; }
  int rc = wait_meta(c->last_ref); if (rc) return rc;
  const size_t total = (size_t) c->last_ref->h_meta[c->last_level].n_total * c->C;
  if (total == 0) { *frac = 0.0f; return BPVO_B200_OK; }
  if (thresh == c->p.goodPointThreshold) {
    // h_mail->lin is the LinOut of the last linearize / estimate_pose
    *frac = (float) c->h_mail->lin.n_good / static_cast<float>(total);
    return BPVO_B200_OK;
  }
  std::vector<float> w(total); size_t cnt = total;
  rc = export_last(c, w.data(), nullptr, &cnt); if (rc) return rc;
  size_t n = 0; for (size_t i = 0; i < cnt; ++i) n += (w[i] > thresh);
  *frac = n / static_cast<float>(cnt);
  return BPVO_B200_OK;
}

// getPointCloudFromRefFrame (vo.cc:249-281) assembled on the device, one D2H
extern "C" int bpvo_b200_point_cloud(bpvo_b200_ctx* c, const bpvo_b200_frame* ref, bpvo_b200_point_info* records, int* n) {
  static_assert(sizeof(bpvo_b200_point_info) == 32 && sizeof(PointInfo) == 32, "32-byte point records (bpvo::PointWithInfo)");
  if (!c || !ref || !n) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  if (ref->ctx != c) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "frame belongs to another ctx");
  if (!ref->has_template) return bp_fail(BPVO_B200_ERR_NO_DATA, "frame has no template");
  if (c->last_ref != ref || c->last_level != c->p.maxTestLevel) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "size mismatch: the last linearize did not run against this frame");
  int rc = wait_meta(ref); if (rc) return rc;
  const int level = c->p.maxTestLevel, np = ref->h_meta[level].n, cap = *n;
  *n = np;
  if (!records || np == 0) return BPVO_B200_OK;
  if (cap < np) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "point cloud buffer too small (%d < %d)", cap, np);
  CUDA_TRY(cudaSetDevice(c->p.device_id));
  // sigma of the last linearize (already on the host after estimate_pose / linearize)
  const float sigma = c->h_mail->lin.sigma;
  const LevelGeom& g = c->geom[level];
  PointInfo* d = reinterpret_cast<PointInfo*>(c->export_buf);          // capacity: capmax * C * 6 floats and at least 8 floats per point
  BP_SWITCH_C(c->C, k_point_cloud<CC><<<ceil_div(np, 256), 256, 0, c->stream>>>(ref->pts[level], np, ref->pyr[0], c->rows, c->cols, u8_pitch(c->cols), g.fx, g.fy, g.cx, g.cy, c->work.res, sigma, c->p.lossFunction, d));
  LAUNCH_CHECK(c);
  CUDA_TRY(cudaMemcpyAsync(records, d, (size_t) np * sizeof(PointInfo), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(bp_sync_stream(c));
  c->counters.d2h_bytes += (int64_t) np * sizeof(PointInfo);
  return BPVO_B200_OK;
}

// Parity hook: `n` consecutive linearize() evaluations of one level THROUGH THE PERSISTENT KERNEL at the caller's poses
// (see DebugArgs in device_types.h).  Residuals / valid flags / weights of the last one are then served by get_residuals & co.
extern "C" int bpvo_b200_debug_device_linearize(bpvo_b200_ctx* c, const bpvo_b200_frame* ref, const bpvo_b200_frame* cur, int level,
                                                 const float* T, int n, bpvo_b200_lin_out* out, int grid_ctas, int cache_bytes) {
  static_assert(sizeof(bpvo_b200_lin_out) == sizeof(LinOut), "bpvo_b200_lin_out mirrors LinOut");
  int rc = check_pair(c, ref, cur); if (rc) return rc;
  if (level < c->p.maxTestLevel || level >= c->L || !T || !out || n <= 0 || n > 1024) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "bad level / poses / count");
  if (!c->coop) return bp_fail(BPVO_B200_ERR_UNSUPPORTED, "device without cooperative launch");
  CUDA_TRY(cudaSetDevice(c->p.device_id));
  M44* d_poses = nullptr; LinOut* d_out = nullptr;
  CUDA_TRY(cudaMalloc(&d_poses, (size_t) n * sizeof(M44)));
  cudaError_t e = cudaMalloc(&d_out, (size_t) n * sizeof(LinOut));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_poses, T, (size_t) n * sizeof(M44), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_out, 0, (size_t) n * sizeof(LinOut), c->stream);
  if (e == cudaSuccess) {
    SolveOverride ov;
    ov.dbg.poses = d_poses; ov.dbg.out = d_out; ov.dbg.n = n; ov.dbg.level = level;
    ov.grid = grid_ctas; ov.cache_bytes = cache_bytes;
    M44 T0; memcpy(T0.m, T, sizeof(T0.m));
    BP_SWITCH_C(c->C, rc = launch_estimate_pose<CC>(c, ref, cur, T0, &ov));
    if (rc == BPVO_B200_OK) {
      e = cudaMemcpyAsync(out, d_out, (size_t) n * sizeof(LinOut), cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(c->h_mail, c->d_mail, sizeof(Mailbox), cudaMemcpyDeviceToHost, c->stream);
    }
  }
  cudaError_t e2 = bp_sync_stream(c);
  cudaFree(d_poses); cudaFree(d_out);
  if (rc) return rc;
  if (e != cudaSuccess || e2 != cudaSuccess) return bp_fail(BPVO_B200_ERR_CUDA, "debug_device_linearize failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
  if (c->h_mail->aborted || c->h_mail->stats[level].status == -4) return bp_fail(BPVO_B200_ERR_CUDA, "on-device GN loop: a grid barrier / exchange timed out");
  if (c->h_mail->stats[level].status == -3) return bp_fail(BPVO_B200_ERR_NO_POINTS, "you should call setData before calling computeResiduals");
  c->last_ref = ref; c->last_level = level;
  return BPVO_B200_OK;
}

// Diagnosis hook: per-linearize trace of the on-device GN loop (level, eval, f_norm, |dp|, max|G|, sigma, scale path, status)
extern "C" int bpvo_b200_debug_set_trace(bpvo_b200_ctx* c, int enable) {
  if (!c) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null ctx");
  CUDA_TRY(cudaSetDevice(c->p.device_id));
  CUDA_TRY(bp_sync_stream(c));
  if (enable && !c->d_trace) {
    CUDA_TRY(cudaMalloc(&c->d_trace, (size_t) kTraceRows * kTraceCols * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c->d_trace_rows, sizeof(int)));
  }
  if (!enable && c->d_trace) { cudaFree(c->d_trace); cudaFree(c->d_trace_rows); c->d_trace = nullptr; c->d_trace_rows = nullptr; }
  if (c->d_trace) CUDA_TRY(cudaMemset(c->d_trace_rows, 0, sizeof(int)));
  return BPVO_B200_OK;
}
extern "C" int bpvo_b200_debug_get_trace(bpvo_b200_ctx* c, float* rows, int max_rows, int* n_rows, int reset) {
  if (!c || !n_rows) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  *n_rows = 0;
  if (!c->d_trace) return BPVO_B200_OK;
  CUDA_TRY(cudaSetDevice(c->p.device_id));
  CUDA_TRY(bp_sync_stream(c));
  int n = 0;
  CUDA_TRY(cudaMemcpy(&n, c->d_trace_rows, sizeof(int), cudaMemcpyDeviceToHost));
  n = std::min(n, kTraceRows);
  *n_rows = n;
  if (rows && max_rows > 0) CUDA_TRY(cudaMemcpy(rows, c->d_trace, (size_t) std::min(n, max_rows) * kTraceCols * sizeof(float), cudaMemcpyDeviceToHost));
  if (reset) CUDA_TRY(cudaMemset(c->d_trace_rows, 0, sizeof(int)));
  return BPVO_B200_OK;
}

// host evaluation of the persistent kernel's per-level shared-memory plan (tests / documentation)
extern "C" int bpvo_b200_debug_cache_plan(int channels, int cache_bytes, int slots_needed, unsigned out[8]) {
  if (!out || (channels != 1 && channels != 3 && channels != 5 && channels != 8)) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "bad argument");
  TplCache t = tpl_cache_off();
  BP_SWITCH_C(channels, t = tpl_cache_plan<CC>((unsigned) kScratchBytes, cache_bytes, slots_needed));
  out[0] = t.pts; out[1] = t.f[TC_I0]; out[2] = t.f[TC_GX]; out[3] = t.f[TC_GY]; out[4] = t.f[TC_R]; out[5] = t.valid;
  out[6] = (unsigned) t.K; out[7] = (unsigned) kScratchBytes;
  return BPVO_B200_OK;
}

// device-time of back-to-back linearize launches (no host round trip), for the roofline figure
extern "C" int bpvo_b200_time_linearize(bpvo_b200_ctx* c, const bpvo_b200_frame* ref, const bpvo_b200_frame* cur, int level,
                                         const float T[16], int iters, int flush_l2, float* ms_per_iter) {
  int rc = check_pair(c, ref, cur); if (rc) return rc;
  if (!T || !ms_per_iter || iters <= 0) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "bad argument");
  CUDA_TRY(cudaSetDevice(c->p.device_id));
  M44 Tm; memcpy(Tm.m, T, sizeof(Tm.m));
  const size_t flush_bytes = (size_t) 256 << 20;
  if (flush_l2 && !c->flush_buf) CUDA_TRY(cudaMalloc(&c->flush_buf, flush_bytes));
  double total = 0;
  for (int i = 0; i < iters; ++i) {
    if (flush_l2) CUDA_TRY(cudaMemsetAsync(c->flush_buf, i & 0xff, flush_bytes, c->stream));
    k_reset_scale<<<1, 1, 0, c->stream>>>(c->work.scale);
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    BP_SWITCH_C(c->C, rc = launch_linearize<CC>(c, ref, cur, level, Tm));
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaEventSynchronize(c->ev1));
    float ms = 0; CUDA_TRY(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    total += ms;
  }
  *ms_per_iter = (float) (total / iters);
  return BPVO_B200_OK;
}
