/*
 * bpvo_b200.h -- C ABI of the B200-native dense-alignment engine that sits under
 * bpvo::VisualOdometry::addFrame() (reference: halismai/bpvo @ 343d9da).
 *
 * The reference has no FFI of its own; its seam is the pair of internal C++ classes that bpvo/vo.cc
 * drives (VisualOdometryFrame, VisualOdometryPoseEstimator) plus the public VisualOdometry class that
 * the programs under apps/ and matlab/vo_mex.cc:204-208 bind.  Every entry point below names the reference
 * interface it replaces (file:line relative to the reference root).  INTEGRATION.md shows the
 * reference-side shim a maintainer would add.
 *
 * Conventions
 *  - plain C, no C++/torch types; all matrices are explicit float arrays in COLUMN-MAJOR order
 *    (Eigen's default storage: pass Matrix44::data() straight through);
 *  - images are row-major, contiguous: uint8 gray `rows x cols`, float32 disparity `rows x cols`;
 *  - every function returns 0 on success, a negative bpvo_b200_status otherwise (the reference throws
 *    bpvo::Error, bpvo/utils.h:211-220; the C++ shim in bpvo_b200/csrc/host re-throws) and never
 *    calls exit(); bpvo_b200_last_error() returns the message of the calling thread's last failure;
 *  - the caller owns every host pointer it passes; the library copies what it keeps
 *    (reference: vo_frame.cc:50-51);
 *  - one ctx = one CUDA device + one stream; a ctx is NOT thread-safe (neither is
 *    bpvo::VisualOdometry), different ctxs may be driven from different threads.
 *  - there is NO CPU fallback: without a CUDA device bpvo_b200_create / bpvo_b200_vo_create fail.
 */
#ifndef BPVO_B200_H
#define BPVO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPVO_B200_VERSION 100
#define BPVO_B200_MAX_LEVELS 16

typedef enum {
  BPVO_B200_OK = 0,
  BPVO_B200_ERR_INVALID_ARG = -1,   /* nullptr image/disparity (vo.cc:68), bad params (dense_descriptor_pyramid.cc:37-38) */
  BPVO_B200_ERR_NO_DATA = -2,       /* "no data in frame" (vo_frame.cc:63) */
  BPVO_B200_ERR_NO_POINTS = -3,     /* "you should call setData before calling computeResiduals" (template_data.cc:177) */
  BPVO_B200_ERR_CUDA = -4,          /* no device / CUDA runtime failure */
  BPVO_B200_ERR_UNSUPPORTED = -5,   /* a DescriptorType / option outside the hot path */
  BPVO_B200_ERR_COMM = -6           /* multi-GPU exchange failure */
} bpvo_b200_status;

/* numeric values are the reference's (bpvo/types.h:125-166, 399-418) */
enum { BPVO_B200_HUBER = 0x10, BPVO_B200_TUKEY = 0x11, BPVO_B200_L2 = 0x12 };
enum { BPVO_B200_INTENSITY = 0x30, BPVO_B200_INTENSITY_AND_GRADIENT = 0x31, BPVO_B200_DESCRIPTOR_FIELDS = 0x32, BPVO_B200_BITPLANES = 0x37 };
enum { BPVO_B200_CD3 = 0, BPVO_B200_CD5 = 1 };
enum { BPVO_B200_LINEAR = 0, BPVO_B200_COSINE = 1, BPVO_B200_CUBIC = 2, BPVO_B200_CUBIC_HERMITE = 3 };   /* InterpolationType, types.h:163-169 */
enum { BPVO_B200_PARAM_TOL = 0x30, BPVO_B200_FUNC_TOL = 0x31, BPVO_B200_GRAD_TOL = 0x32,
       BPVO_B200_MAX_ITERS = 0x33, BPVO_B200_SOLVER_ERROR = 0x34 };
enum { BPVO_B200_KF_LARGE_TRANSLATION = 0x40, BPVO_B200_KF_LARGE_ROTATION = 0x41,
       BPVO_B200_KF_SMALL_FRAC_GOOD = 0x42, BPVO_B200_KF_NONE = 0x43, BPVO_B200_KF_FIRST_FRAME = 0x44 };

/* flags */
#define BPVO_B200_FLAG_NO_GRAPHS    2  /* launch the per-frame kernel sequences one by one instead of as CUDA graphs */
#define BPVO_B200_FLAG_FAST_BLEND   4  /* bit-planes: evaluate the 4-tap bilinear blend of the residuals with fp32 FMAs on fp64-derived
                                          fractions instead of the reference's double expression (photo_error.cc:381-388).  |dr| <= 2e-7
                                          (north star: 1e-5) but the residual vector is no longer bit-identical to the reference, and the
                                          same-box A/B gained only 2 % (profiles/README.md): off by default.  Projection, Floor and validity
                                          are fp64 in both modes; intensity always uses the double expression. */
#define BPVO_B200_FLAG_TMA_DESCRIPTOR 8 /* bit-planes descriptor kernel with its tiles on the TMA engine (cp.async.bulk.tensor load of the u8 halo tile,
                                          cp.async.bulk.tensor store of the 8 x 64 x 8 f32 result tile): same bits, 14 % slower than the default
                                          kernel on B200 (one tile per CTA cannot hide the TMA round trips): off by default */
#define BPVO_B200_FLAG_HOST_SOLVE   1  /* estimate_pose drives the GN loop from the host (one sync per iteration)
                                          instead of the on-device loop; same results, used for parity tests */

/* POD mirror of the hot-path fields of bpvo::AlgorithmParameters (bpvo/types.h:171-397,
 * defaults bpvo/types.cc:31-66 via bpvo_b200_default_params). */
typedef struct {
  int32_t numPyramidLevels;               /* <=0: auto (vo.cc:101-104) */
  int32_t minImageDimensionForPyramid;
  float   sigmaPriorToCensusTransform;    /* > 0: cv::GaussianBlur(3x3) on u8 before the census (census.cc:63-65), OpenCV-4 fixed point */
  float   sigmaBitPlanes;
  int32_t maxIterations;
  float   parameterTolerance;
  float   functionTolerance;
  float   gradientTolerance;
  int32_t relaxTolerancesForCoarseLevels; /* unused by the reference too (no reader) */
  int32_t gradientEstimation;             /* BPVO_B200_CD3 / CD5 (template_data.cc:117-130) */
  int32_t interp;                         /* only kLinear (photo_error.cc:381-389) */
  int32_t lossFunction;
  int32_t descriptor;                     /* kIntensity (1 channel), kIntensityAndGradient (3), kDescriptorFieldsFirstOrder (5), kBitPlanes (8) */
  int32_t verbosity;                      /* ignored: the engine never prints */
  float   minTranslationMagToKeyFrame;
  float   minRotationMagToKeyFrame;
  float   maxFractionOfGoodPointsToKeyFrame;
  float   goodPointThreshold;
  int32_t minNumPixelsForNonMaximaSuppression;
  int32_t nonMaxSuppRadius;
  int32_t minNumPixelsToWork;             /* unused by the reference too */
  float   minSaliency;
  float   minValidDisparity;
  float   maxValidDisparity;
  int32_t maxTestLevel;
  int32_t withNormalization;
  /* engine options (not in the reference) */
  int32_t device_id;                      /* CUDA device ordinal */
  int32_t flags;                          /* BPVO_B200_FLAG_* */
  /* DescriptorFields (bpvo/types.h, defaults 0.75 / 1.75 at types.cc:36-37): smoothing before / after the gradient split */
  float   dfSigma1;
  float   dfSigma2;
} bpvo_b200_params;

/* bpvo::OptimizerStatistics (bpvo/types.h:444-482) */
typedef struct {
  int32_t numIterations;
  float   finalError;
  float   firstOrderOptimality;
  int32_t status;
} bpvo_b200_stats;

/* bpvo::Result (bpvo/types.h:496-566) minus the point cloud, which is fetched separately */
typedef struct {
  float   pose[16];                       /* column-major */
  int32_t isKeyFrame;
  int32_t keyFramingReason;
  int32_t numLevels;
  bpvo_b200_stats optimizerStatistics[BPVO_B200_MAX_LEVELS];
  int32_t numFunEvals;                    /* linearize() calls (= GN iterations) inside this addFrame */
  int32_t numPointCloud;                  /* size of Result::pointCloud, 0 if none */
} bpvo_b200_result;

/* per-phase device-time / launch counters (cudaEvent based; enable with bpvo_b200_set_profiling) */
typedef struct {
  double  ms_upload, ms_pyramid, ms_descriptor, ms_template, ms_linearize, ms_total;
  int64_t launches;                       /* kernels launched by this ctx since creation / reset */
  int64_t linearize_calls;                /* GN iterations (linearize evaluations) */
  int64_t h2d_bytes, d2h_bytes;
  int64_t solve_calls;                    /* estimate_pose calls (= launches of the persistent solve kernel) */
} bpvo_b200_counters;

typedef struct bpvo_b200_ctx   bpvo_b200_ctx;    /* VisualOdometryPoseEstimator + device/stream binding */
typedef struct bpvo_b200_frame bpvo_b200_frame;  /* VisualOdometryFrame */
typedef struct bpvo_b200_vo    bpvo_b200_vo;     /* VisualOdometry */

int  bpvo_b200_version(void);
const char* bpvo_b200_last_error(void);
/* AlgorithmParameters::AlgorithmParameters() (bpvo/types.cc:31-66) */
void bpvo_b200_default_params(bpvo_b200_params* p);
/* number of CUDA devices visible (0 => every create fails with BPVO_B200_ERR_CUDA) */
int  bpvo_b200_device_count(void);

/* ---------------------------------------------------------------------------------------------
 * VisualOdometry level -- what apps/vo_perf.cc:84-87 and matlab/vo_mex.cc:204-208 bind
 * ------------------------------------------------------------------------------------------- */
/* VisualOdometry(const Matrix33& K, float baseline, ImageSize, const AlgorithmParameters&)  (bpvo/vo.h:42-43, vo.cc:94-110) */
int bpvo_b200_vo_create(bpvo_b200_vo** out, const float K[9], float baseline, int rows, int cols, const bpvo_b200_params* p);
int bpvo_b200_vo_destroy(bpvo_b200_vo* vo);
/* Result addFrame(const uint8_t* image, const float* disparity)  (bpvo/vo.h:75, vo.cc:66-72, 125-197) */
int bpvo_b200_vo_add_frame(bpvo_b200_vo* vo, const uint8_t* image, const float* disparity, bpvo_b200_result* result);
/* int numPointsAtLevel(int level = -1) const  (bpvo/vo.h:83, vo.cc:226-238) */
int bpvo_b200_vo_num_points_at_level(const bpvo_b200_vo* vo, int level, int* n);
/* const PointVector& pointsAtLevel(int level = -1) const  (bpvo/vo.h:88, vo.cc:240-247): xyzw, 4*n floats */
int bpvo_b200_vo_points_at_level(const bpvo_b200_vo* vo, int level, float* xyzw, int max_points);
/* const Trajectory& trajectory() const  (bpvo/vo.h:93, trajectory.cc:42-50): 16 floats per pose; returns count in *n */
int bpvo_b200_vo_trajectory(const bpvo_b200_vo* vo, float* poses, int max_poses, int* n);
/* Result::pointCloud of the last addFrame (vo.cc:260-281): xyzw / weight / gray per point */
int bpvo_b200_vo_point_cloud(const bpvo_b200_vo* vo, float* xyzw, float* weights, uint8_t* gray, int max_points, int* n);
/* the engine ctx / reference frame behind a vo (for counters, parity dumps) */
bpvo_b200_ctx* bpvo_b200_vo_ctx(bpvo_b200_vo* vo);
const bpvo_b200_frame* bpvo_b200_vo_ref_frame(const bpvo_b200_vo* vo);

/* ---------------------------------------------------------------------------------------------
 * seam level -- what bpvo/vo.cc calls on VisualOdometryFrame / VisualOdometryPoseEstimator
 * ------------------------------------------------------------------------------------------- */
/* VisualOdometryPoseEstimator(const AlgorithmParameters&)  (vo_pose_estimator.cc:55-59) + device binding.
 * numPyramidLevels must already be resolved (> 0), as at vo.cc:101-104. */
int bpvo_b200_create(bpvo_b200_ctx** out, const float K[9], float baseline, int rows, int cols, const bpvo_b200_params* p);
int bpvo_b200_destroy(bpvo_b200_ctx* ctx);
/* VisualOdometryFrame(const Matrix33& K, float b, const AlgorithmParameters&)  (vo_frame.cc:13-30) */
int bpvo_b200_frame_create(bpvo_b200_ctx* ctx, bpvo_b200_frame** out);
int bpvo_b200_frame_destroy(bpvo_b200_frame* f);
/* void setData(const cv::Mat& image, const cv::Mat& disparity)  (vo_frame.h:39, vo_frame.cc:48-55):
 * copies both, builds the image pyramid and the dense descriptors of levels >= maxTestLevel.
 * Pageable pointers are staged; pinned (cudaHostAlloc/Register'd) host pointers and device pointers are
 * DMA'd directly and stay borrowed until the next synchronising call on this ctx (bpvo_b200_vo_add_frame
 * always synchronises before it returns). */
int bpvo_b200_frame_set_data(bpvo_b200_frame* f, const uint8_t* image, const float* disparity);
/* void setTemplate()  (vo_frame.h:48, vo_frame.cc:61-93) -> TemplateData::setData (template_data.cc:37-142) */
int bpvo_b200_frame_set_template(bpvo_b200_frame* f);
/* hasTemplate() / empty() / clear()  (vo_frame.h:50-53) */
int bpvo_b200_frame_has_template(const bpvo_b200_frame* f);
int bpvo_b200_frame_empty(const bpvo_b200_frame* f);
int bpvo_b200_frame_clear(bpvo_b200_frame* f);
/* numLevels()  (vo_frame.cc:46) and per-level image size */
int bpvo_b200_frame_num_levels(const bpvo_b200_frame* f);
int bpvo_b200_frame_level_size(const bpvo_b200_frame* f, int level, int* rows, int* cols);
/* getTemplateDataAtLevel(l)->numPoints() / points()  (template_data.h:68-71) */
int bpvo_b200_frame_num_points(const bpvo_b200_frame* f, int level, int* n);
int bpvo_b200_frame_get_points(const bpvo_b200_frame* f, int level, float* xyzw /* 4n */);
/* parity dumps in the REFERENCE's layouts (device layouts differ):
 *   pyramid level (ImagePyramid::operator[], image_pyramid.h)          u8  rows x cols
 *   descriptor   (DenseDescriptor::getChannel(c), dense_descriptor.h)  f32 planar channels x rows x cols
 *   saliency     (DenseDescriptor::computeSaliencyMap)                 f32 rows x cols (of the last set_template)
 *   pixels       (TemplateData::pixels())                              f32 channel-major C*N
 *   jacobians    (TemplateData::jacobians())                           f32 channel-major C*N x 6
 *   point_inds   (valid_inds of template_data.cc:69-83)                i32 N  (y*cols + x)
 *   normalization (RigidBodyWarp::_T, warps.cc:27-48)                  f32 4x4 column-major */
int bpvo_b200_frame_get_pyramid(const bpvo_b200_frame* f, int level, uint8_t* out);
int bpvo_b200_frame_get_descriptor(const bpvo_b200_frame* f, int level, float* planes, int* channels);
int bpvo_b200_frame_get_saliency(const bpvo_b200_frame* f, int level, float* out);
int bpvo_b200_frame_get_pixels(const bpvo_b200_frame* f, int level, float* out);
int bpvo_b200_frame_get_jacobians(const bpvo_b200_frame* f, int level, float* out);
int bpvo_b200_frame_get_point_inds(const bpvo_b200_frame* f, int level, int32_t* out);
int bpvo_b200_frame_get_normalization(const bpvo_b200_frame* f, int level, float Tn[16]);

/* fine seam: one PoseEstimatorGN::linearize (pose_estimator_gn.h:70-81) = residuals -> robust scale ->
 * weights -> normal equations.  The host keeps solve() / convergence (pose_estimator_base.h:324-407).
 * first_call_of_level != 0 resets the scale-estimator state (reset(), pose_estimator_base.h:287-293). */
int bpvo_b200_linearize(bpvo_b200_ctx* ctx, const bpvo_b200_frame* ref, const bpvo_b200_frame* cur, int level,
                        const float T[16], int first_call_of_level,
                        float H[36], float G[6], float* f_norm, float* sigma, int* n_valid);
/* coarse seam: std::vector<OptimizerStatistics> estimatePose(ref, cur, T_init, T_est)
 * (vo_pose_estimator.h:48-52, vo_pose_estimator.cc:63-93); stats has numLevels entries; *num_fun_evals
 * (optional) receives the number of linearize() evaluations. */
int bpvo_b200_estimate_pose(bpvo_b200_ctx* ctx, const bpvo_b200_frame* ref, const bpvo_b200_frame* cur,
                            const float T_init[16], float T_est[16], bpvo_b200_stats* stats, int* num_fun_evals);
/* const WeightsVector& getWeights()  (vo_pose_estimator.cc:95-99): C*N weights of the last linearize,
 * channel-major.  *count: in = capacity of `w` in floats (0 = everything; ignored when w == NULL), out = C*N.
 * min(capacity, C*N) entries are copied -- the first N are channel 0, all that vo.cc:264 reads.
 * Same for residuals / valid flags (PoseEstimatorBase::residuals()/getValidFlags(), pose_estimator_base.h:189-223). */
int bpvo_b200_get_weights(bpvo_b200_ctx* ctx, float* w, size_t* count);
int bpvo_b200_get_residuals(bpvo_b200_ctx* ctx, float* r, size_t* count);
int bpvo_b200_get_valid(bpvo_b200_ctx* ctx, uint8_t* v, size_t* count);   /* per point, N entries */
/* float getFractionOfGoodPoints(float thresh)  (vo_pose_estimator.cc:101-107), counted on the device */
int bpvo_b200_fraction_good(bpvo_b200_ctx* ctx, float thresh, float* frac);
/* getPointCloudFromRefFrame (vo.cc:249-281) assembled on the device: per template point of `ref` at maxTestLevel one
 * 32-byte record {x, y, z, w (float), r, g, b, a (uint8), weight (float), 8 bytes of padding} = the memory layout of
 * bpvo::PointWithInfo (bpvo/point_cloud.h:50-58: 16 + 4 + 4 bytes padded to 32, 32-byte stride), so the records can be
 * copied straight into a PointWithInfoVector: colour = the ref frame's full-resolution image at the projection K_l * X (bounds and truncation as
 * vo.cc:269-272), weight = channel 0 of the last linearize's weights (Q7; invalid points carry 1, Q6).  One D2H of
 * 32 N bytes instead of points + weights + the whole image.  *n: in = capacity in records, out = N. */
typedef struct { float x, y, z, w; uint8_t rgba[4]; float weight; uint8_t pad[8]; } bpvo_b200_point_info;   /* 32 bytes */
int bpvo_b200_point_cloud(bpvo_b200_ctx* ctx, const bpvo_b200_frame* ref, bpvo_b200_point_info* records, int* n);

/* ---------------------------------------------------------------------------------------------
 * multi-GPU: template points sharded across ranks, 28-scalar exchange per GN iteration
 * (no reference counterpart: the reference is single-process)
 * ------------------------------------------------------------------------------------------- */
/* 128-byte rendezvous token (an ncclUniqueId) created on rank 0 and distributed by the caller (e.g. torch.distributed,
 * MPI, a file).  NCCL is bound at run time (libnccl.so.2); BPVO_B200_ERR_COMM if it cannot be loaded. */
int bpvo_b200_comm_unique_id(uint8_t id[128]);
/* join (collective over all ranks): afterwards set_template keeps this rank's contiguous scan-order block of points
 * (multiples of 16; every rank must feed the SAME frames), the Hartley sums, the radix-select histograms of the exact
 * median and the 30 fp64 normal-equation sums are all-reduced over NCCL, and estimate_pose runs the host-driven loop:
 * every rank computes bit-identical H, G, sigma and therefore identical poses and key-frame decisions.
 * frame_num_points / get_points / get_weights / get_residuals then refer to the local shard. */
int bpvo_b200_comm_init(bpvo_b200_ctx* ctx, int rank, int nranks, const uint8_t id[128]);
int bpvo_b200_comm_destroy(bpvo_b200_ctx* ctx);
/* Peer-memory mode (optional, after comm_init; one process per GPU on one NVLink/NVSwitch node, <= 8 ranks): the
 * point-sharded estimate_pose runs the SAME persistent on-device GN loop as a single GPU, and the ranks exchange the
 * bracket histogram + the handful of median candidates + the 30 fp64 sums INSIDE the kernel through flag-in-data
 * mailboxes in each other's memory (CUDA IPC mappings, stores over NVLink) -- no NCCL launch and no host round trip
 * per iteration.  peer_export allocates this rank's mailbox and returns its cudaIpcMemHandle_t (64 bytes, distribute
 * to all ranks e.g. with torch.distributed.all_gather); peer_init maps the peers' mailboxes (handles in rank order). */
int bpvo_b200_peer_export(bpvo_b200_ctx* ctx, uint8_t handle[64]);
int bpvo_b200_peer_init(bpvo_b200_ctx* ctx, const uint8_t* handles /* nranks x 64 bytes */);
/* pyramid levels whose template has fewer points than this are REPLICATED (every rank keeps all points and runs the
 * level like a single GPU, bit-identically, no exchange) instead of sharded: "when the point count justifies it".
 * Default 131072; 0 shards every level.  Applies to templates built afterwards. */
int bpvo_b200_peer_set_min_points(bpvo_b200_ctx* ctx, int min_points);

/* ---------------------------------------------------------------------------------------------
 * measurement helpers
 * ------------------------------------------------------------------------------------------- */
/* throughput mode: SMs (one CTA each) the on-device GN loop of this ctx uses, 0 = all.  With e.g. 4 ctxs of 37 CTAs, four
 * independent VisualOdometry streams (one host thread each) run their solves concurrently on one GPU. */
int bpvo_b200_set_solver_ctas(bpvo_b200_ctx* ctx, int ctas);
int bpvo_b200_set_profiling(bpvo_b200_ctx* ctx, int enable);   /* cudaEvent pairs around each phase */
int bpvo_b200_get_counters(bpvo_b200_ctx* ctx, bpvo_b200_counters* out);
/* SM-cycle counters of the phases of the on-device GN loop (CTA 0), accumulated while profiling is on:
 * P1, sync, P2, sync, P3, sync, scale, P4, sync, final-sum, solve, other, ... */
int bpvo_b200_get_phase_cycles(bpvo_b200_ctx* ctx, long long cycles[64], int reset);
int bpvo_b200_reset_counters(bpvo_b200_ctx* ctx);
int bpvo_b200_synchronize(bpvo_b200_ctx* ctx);
/* cudaEvent pair on the ctx stream: start records an event, stop records another, waits for it and returns
 * the device time between them (what bench.py times its steps with) */
int bpvo_b200_timer_start(bpvo_b200_ctx* ctx);
int bpvo_b200_timer_stop(bpvo_b200_ctx* ctx, float* ms);
/* linearize() evaluations per pyramid level of the last estimate_pose (numLevels ints) */
int bpvo_b200_last_level_evals(bpvo_b200_ctx* ctx, int* evals);
/* device time (microseconds, GPU global timer) the on-device GN loop of the last estimate_pose spent in each pyramid level */
int bpvo_b200_last_level_us(bpvo_b200_ctx* ctx, float* us);
/* the phase counters of get_phase_cycles split by pyramid level: [BPVO_B200_MAX_LEVELS][16] */
int bpvo_b200_get_level_phase_cycles(bpvo_b200_ctx* ctx, long long* cycles, int reset);
/* Parity hooks of the on-device GN loop (the kernel bpvo_b200_estimate_pose launches; no reference counterpart).
 * debug_device_linearize: `n` consecutive PoseEstimatorGN::linearize evaluations (pose_estimator_gn.h:70-81) of `level`
 * executed INSIDE that persistent kernel at the caller's poses T[0..n) (16 floats each, column-major) -- its shared-memory
 * template cache, its bracketed exact median (from the 2nd evaluation on), its flag-in-data exchange of the sums -- with no
 * solve and no pose update; the scale-estimator state is reset before the first.  out[e] must equal what e + 1 calls of
 * bpvo_b200_linearize (first_call_of_level on the first) return; get_residuals / get_valid / get_weights then serve the
 * last evaluation.  grid_ctas (0 = one CTA per SM) and cache_bytes (-1 = all shared memory) force the multi-slot and
 * partially-cached code paths at small sizes.  scale_path: 0 = scale kept, 1 = radix select, 3 = bracketed select. */
typedef struct { float H[36]; float G[6]; float f_norm; float sigma; int32_t n_valid; int32_t n_good; int32_t solve_ok; int32_t scale_path; } bpvo_b200_lin_out;
int bpvo_b200_debug_device_linearize(bpvo_b200_ctx* ctx, const bpvo_b200_frame* ref, const bpvo_b200_frame* cur, int level,
                                     const float* T, int n, bpvo_b200_lin_out* out, int grid_ctas, int cache_bytes);
/* per-linearize trace of the GN loop inside bpvo_b200_estimate_pose: rows of 8 floats {level, evaluation, f_norm, |dp|,
 * max|G|, sigma, scale_path, status so far}, what PoseEstimatorBase::run prints at verbosity kIteration
 * (pose_estimator_base.h:231-247).  set_trace(1) allocates and clears the buffer (8192 rows), set_trace(0) frees it. */
int bpvo_b200_debug_set_trace(bpvo_b200_ctx* ctx, int enable);
int bpvo_b200_debug_get_trace(bpvo_b200_ctx* ctx, float* rows, int max_rows, int* n_rows, int reset);
/* the on-device GN loop's shared-memory plan for a level whose threads own `slots_needed` points each, given
 * `cache_bytes` of dynamic shared memory: byte offsets of {points, I0, gx, gy, residuals, valid flags}
 * (0xffffffff = that field stays in global memory), slots, offset of the cache area.  Host-side evaluation of the
 * kernel's own planning function, for tests and documentation. */
int bpvo_b200_debug_cache_plan(int channels, int cache_bytes, int slots_needed, unsigned out[8]);
/* pinned host allocation for zero-staging uploads (cudaHostAlloc / cudaFreeHost) */
void* bpvo_b200_host_alloc(size_t bytes);
void  bpvo_b200_host_free(void* p);
/* time `iters` back-to-back linearize() launches at a fixed pose on the ctx stream with cudaEvents
 * (device time only, no host round trip per iteration); returns the average ms per linearize */
int bpvo_b200_time_linearize(bpvo_b200_ctx* ctx, const bpvo_b200_frame* ref, const bpvo_b200_frame* cur, int level,
                             const float T[16], int iters, int flush_l2, float* ms_per_iter);

/* =============================================================================================
 * Upstream stereo: the disparity producer in front of the path.
 * Replaces bpvo::StereoAlgorithm (utils/stereo_algorithm.h:14-37) for its default algorithm "BlockMatching":
 * utils/stereo_algorithm.cc:67-85 fills OpenCV's CvStereoBMState from the config file (the fields below, same names and
 * defaults), :99-111 runs cvFindStereoCorrespondenceBM and converts CV_16S -> CV_32F with 1/16.  Results are bit-identical to
 * OpenCV's StereoBM (pinned against cv2 4.13, see oracle/stereo_oracle.cc).  Not accelerated, rejected at creation with
 * BPVO_B200_ERR_UNSUPPORTED: the NORMALIZED_RESPONSE pre-filter, minDisparity > 0, speckle filtering, the left-right check
 * (none of them is used by the reference's configurations), SADWindowSize > 31, numberOfDisparities > 256; the other
 * algorithms of StereoAlgorithm (SGBM, and the GPL-gated SGM / RSGM) stay on the CPU.
 * ============================================================================================= */
enum { BPVO_B200_STEREO_BM_NORMALIZED_RESPONSE = 0, BPVO_B200_STEREO_BM_XSOBEL = 1 };   /* CV_STEREO_BM_* */
typedef struct bpvo_b200_stereo bpvo_b200_stereo;
typedef struct {
  int32_t numberOfDisparities;   /* no default: "must be provided" (stereo_algorithm.cc:75); multiple of 16 */
  int32_t SADWindowSize;         /* 15 */
  int32_t minDisparity;          /* 0 */
  int32_t preFilterType;         /* CV_STEREO_BM_XSOBEL */
  int32_t preFilterSize;         /* 9 (unused by XSOBEL) */
  int32_t preFilterCap;          /* 31 */
  int32_t textureThreshold;      /* 10 */
  int32_t uniquenessRatio;       /* 15 */
  int32_t speckleWindowSize;     /* 0 */
  int32_t speckleRange;          /* 0 */
  int32_t trySmallerWindows;     /* read from the config (conf/kitti.cfg:13) and ignored, as OpenCV ignores it */
  int32_t disp12MaxDiff;         /* -1 */
  int32_t device_id;
} bpvo_b200_stereo_params;
void bpvo_b200_stereo_default_params(bpvo_b200_stereo_params* p);
/* StereoAlgorithm::StereoAlgorithm (stereo_algorithm.cc:156-160); OpenCV's argument checks with OpenCV's messages */
int bpvo_b200_stereo_create(bpvo_b200_stereo** out, int rows, int cols, const bpvo_b200_stereo_params* p);
int bpvo_b200_stereo_destroy(bpvo_b200_stereo* s);
/* StereoAlgorithm::run (stereo_algorithm.cc:163-166).  left / right: rows x cols u8, row-major, host or device memory.
 * dmap: rows x cols f32 disparities in pixels (invalid = minDisparity - 1); disp16: OpenCV's CV_16S map (4 fractional bits);
 * host or device memory, either may be NULL.  Returns when the results are in place.  The work runs on the object's own
 * stream: device-resident inputs must be complete when the call is made (synchronize the producing stream first). */
int bpvo_b200_stereo_run(bpvo_b200_stereo* s, const uint8_t* left, const uint8_t* right, float* dmap, int16_t* disp16);
/* StereoAlgorithm::getInvalidValue (stereo_algorithm.cc:138-146, 168) with the reference's arithmetic: short(minDisparity - 1) / 16.0f
 * (-0.0625 for minDisparity 0) -- not the value invalid pixels carry in dmap, which is filtered_value = minDisparity - 1 */
float bpvo_b200_stereo_invalid_value(const bpvo_b200_stereo* s);
float bpvo_b200_stereo_filtered_value(const bpvo_b200_stereo* s);
/* parity dump: the XSOBEL pre-filtered pair of the last run (rows x cols u8 each, host memory, either may be NULL) */
int bpvo_b200_stereo_get_prefiltered(bpvo_b200_stereo* s, uint8_t* left, uint8_t* right);
/* device time of the last run's kernels (CUDA events on the object's stream); kernels launched so far */
int bpvo_b200_stereo_last_kernel_ms(bpvo_b200_stereo* s, float* ms);
long long bpvo_b200_stereo_launches(const bpvo_b200_stereo* s);
/* image pair in, pose out (utils/dataset.cc:133 feeding VisualOdometry::addFrame): the disparity map stays on the device */
int bpvo_b200_vo_add_stereo_frame(bpvo_b200_vo* vo, bpvo_b200_stereo* s, const uint8_t* left, const uint8_t* right, bpvo_b200_result* result);

#ifdef __cplusplus
}
#endif
#endif /* BPVO_B200_H */
