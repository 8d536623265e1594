/*
 * bpvo_oracle.h -- C API of the CPU restatement ("oracle") of halismai/bpvo's per-frame
 * Gauss-Newton dense-alignment path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product (bpvo_b200/) never
 * links, imports or calls anything in oracle/.
 *
 * PARITY STATUS: pinned BIT-EXACT against the reference's own sources compiled from /root/reference
 * against stand-in Eigen/OpenCV headers (oracle/_ref, tests/test_oracle_vs_reference.py): every stage
 * up to whole addFrame streams (poses, per-level iteration counts, key-frame decisions).  The reference's
 * own tests pin nothing on this path (no golden vectors / assertions); the third-party
 * image ops (cv::pyrDown, cv::GaussianBlur) are pinned against cv2 4.13 golden vectors in tests/golden/.
 *
 * All matrices cross this API column-major (Eigen's default storage).
 * Every function cites the reference file:line it restates in bpvo_oracle.cc.
 */
#ifndef BPVO_ORACLE_H
#define BPVO_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* enum values are the reference's (bpvo/types.h:125-166, 399-420) */
enum { ORC_HUBER = 0x10, ORC_TUKEY = 0x11, ORC_L2 = 0x12 };
enum { ORC_INTENSITY = 0x30, ORC_INTENSITY_AND_GRADIENT = 0x31, ORC_DESCRIPTOR_FIELDS = 0x32, ORC_BITPLANES = 0x37 };
enum { ORC_CD3 = 0, ORC_CD5 = 1 };
enum { ORC_LINEAR = 0, ORC_COSINE = 1, ORC_CUBIC = 2, ORC_CUBIC_HERMITE = 3 };
enum { ORC_PARAM_TOL = 0x30, ORC_FUNC_TOL, ORC_GRAD_TOL, ORC_MAX_ITERS, ORC_SOLVER_ERROR };
enum { ORC_KF_LARGE_TRANSLATION = 0x40, ORC_KF_LARGE_ROTATION, ORC_KF_SMALL_FRAC_GOOD,
       ORC_KF_NONE, ORC_KF_FIRST_FRAME };

/* POD mirror of bpvo::AlgorithmParameters hot-path fields (bpvo/types.h:171-397), same field
 * order and layout as bpvo_b200_params in include/bpvo_b200.h, plus oracle-only switches. */
typedef struct {
  int32_t numPyramidLevels;
  int32_t minImageDimensionForPyramid;
  float   sigmaPriorToCensusTransform;
  float   sigmaBitPlanes;
  int32_t maxIterations;
  float   parameterTolerance;
  float   functionTolerance;
  float   gradientTolerance;
  int32_t relaxTolerancesForCoarseLevels;
  int32_t gradientEstimation;
  int32_t interp;
  int32_t lossFunction;
  int32_t descriptor;
  int32_t verbosity;
  float   minTranslationMagToKeyFrame;
  float   minRotationMagToKeyFrame;
  float   maxFractionOfGoodPointsToKeyFrame;
  float   goodPointThreshold;
  int32_t minNumPixelsForNonMaximaSuppression;
  int32_t nonMaxSuppRadius;
  int32_t minNumPixelsToWork;
  float   minSaliency;
  float   minValidDisparity;
  float   maxValidDisparity;
  int32_t maxTestLevel;
  int32_t withNormalization;
  /* oracle-only */
  int32_t use_rcp;      /* 1 = _mm_rcp_ps Jacobians as the reference (rigid_body_warp.cc:47-58) */
  int32_t num_threads;  /* 1 = reference default build (WITH_TBB off); >1 = OpenMP stand-in   */
  float   dfSigma1;     /* DescriptorFields: smoothing before / after the gradient split (types.h, defaults 0.75 / 1.75) */
  float   dfSigma2;
} orc_params;

typedef struct {
  int32_t numIterations;
  float   finalError;
  float   firstOrderOptimality;
  int32_t status;
} orc_stats;

typedef struct {
  float   pose[16];          /* column-major 4x4 */
  int32_t isKeyFrame;
  int32_t keyFramingReason;
  int32_t numLevels;
  orc_stats stats[16];
  int32_t numFunEvals;       /* total linearize() calls inside this addFrame (GN iterations) */
  int32_t numPointCloud;     /* points in Result::pointCloud (0 if none) */
} orc_result;

void orc_default_params(orc_params* p);   /* bpvo/types.cc:31-66 */

/* ---- stage-level functions (unit parity) ---- */
/* cv::pyrDown on u8 (bpvo/image_pyramid.cc:49); dst is ((rows+1)/2) x ((cols+1)/2) */
void orc_pyr_down(const uint8_t* src, int rows, int cols, uint8_t* dst);
/* cv::GaussianBlur 5x5 f32 reflect-101 (bpvo/bitplanes_descriptor.cc:56) */
void orc_gaussian_blur5(const float* src, int rows, int cols, float sigma, float* dst);
/* cv::GaussianBlur on CV_32F with any odd kernel size; ksize <= 0: derived from sigma like cv::Size() (gradient_descriptor.cc:52-53) */
void orc_gaussian_blur_f32(const float* src, int rows, int cols, int ksize, double sigma, float* dst);
/* bpvo/census.cc:59-91 with sigma<=0 */
void orc_census(const uint8_t* src, int rows, int cols, uint8_t* dst);
void orc_gaussian_blur3_u8(const uint8_t* src, int rows, int cols, float sigma, uint8_t* dst);   /* cv::GaussianBlur 3x3 on CV_8U, cv2-4.13 fixed point */
/* descriptor of one level: planar C x rows x cols f32 (bitplanes_descriptor.cc:84-91 / intensity_descriptor.cc:31-43) */
int  orc_descriptor(const orc_params* p, const uint8_t* img, int rows, int cols, float* planes);
/* DenseDescriptor::computeSaliencyMap incl. its live indexing bugs (dense_descriptor.cc:92-100, imgproc.cc:45-142) */
void orc_saliency(const float* planes, int channels, int rows, int cols, float* dst);
/* exact median as bpvo/utils.h:224-252 (modifies buf) */
float orc_median(float* buf, size_t n);
/* 6x6 solve as PoseEstimatorData_::solve (pose_estimator_base.h:90-148); returns ok flag */
int  orc_solve6(const float H[36], const float G[6], float dp[6]);
/* paramsToPose: Tn^-1 * exp(p) * Tn (rigid_body_warp.h:130-138, math_utils.h:140-168) */
void orc_params_to_pose(const float Tn[16], const float p[6], float out[16]);

/* ---- frame / pose-estimator level (the seam vo.cc drives) ---- */
typedef struct orc_frame orc_frame;
typedef struct orc_estimator orc_estimator;

orc_frame* orc_frame_create(const float K[9], float baseline, int rows, int cols, const orc_params* p);
void orc_frame_destroy(orc_frame*);
void orc_frame_set_data(orc_frame*, const uint8_t* image, const float* disparity);   /* vo_frame.cc:48-55 */
int  orc_frame_set_template(orc_frame*);                                             /* vo_frame.cc:61-93 */
int  orc_frame_num_levels(const orc_frame*);
void orc_frame_level_size(const orc_frame*, int level, int* rows, int* cols);
const uint8_t* orc_frame_pyramid(const orc_frame*, int level);
const float* orc_frame_descriptor(const orc_frame*, int level, int* channels);       /* planar */
const float* orc_frame_saliency(const orc_frame*, int level);                        /* of last set_template */
int  orc_frame_num_points(const orc_frame*, int level);
const float* orc_frame_points(const orc_frame*, int level);        /* N x 4 */
const float* orc_frame_pixels(const orc_frame*, int level);        /* C*N channel-major */
const float* orc_frame_jacobians(const orc_frame*, int level);     /* (C*N+1) x 6 channel-major */
const int32_t* orc_frame_point_inds(const orc_frame*, int level);  /* y*cols+x of each point */
void orc_frame_normalization(const orc_frame*, int level, float Tn[16]);

orc_estimator* orc_estimator_create(const orc_params* p);
void orc_estimator_destroy(orc_estimator*);
/* one PoseEstimatorGN::linearize (pose_estimator_gn.h:70-81). reset_scale!=0 mimics reset() at run() start */
float orc_linearize(orc_estimator*, const orc_frame* ref, const orc_frame* cur, int level,
                    const float T[16], int reset_scale, float H[36], float G[6], float* sigma);
size_t orc_estimator_num_residuals(const orc_estimator*);
const float* orc_estimator_residuals(const orc_estimator*);
const float* orc_estimator_weights(const orc_estimator*);
const uint16_t* orc_estimator_valid(const orc_estimator*);          /* replicated, C*N */
/* VisualOdometryPoseEstimator::estimatePose (vo_pose_estimator.cc:63-93); returns #linearize calls */
int  orc_estimate_pose(orc_estimator*, const orc_frame* ref, const orc_frame* cur,
                       const float T_init[16], float T_est[16], orc_stats* stats /*numLevels*/);
float orc_fraction_good(const orc_estimator*, float thresh);        /* vo_pose_estimator.cc:101-107 */
/* per-iteration trace of run(): rows of 8 floats {level, eval, f_norm, |dp|, max|G|, sigma, converged, status} = the table
 * PoseEstimatorBase::run prints at verbosity kIteration (pose_estimator_base.h:231-247, 295-299) */
void orc_estimator_set_trace(orc_estimator*, int on);
int  orc_estimator_get_trace(orc_estimator*, float* rows, int max_rows);

/* ---- VisualOdometry level (bpvo/vo.cc:94-281) ---- */
typedef struct orc_vo orc_vo;
orc_vo* orc_vo_create(const float K[9], float baseline, int rows, int cols, const orc_params* p);
void orc_vo_destroy(orc_vo*);
int  orc_vo_add_frame(orc_vo*, const uint8_t* image, const float* disparity, orc_result* out);
int  orc_vo_num_points_at_level(const orc_vo*, int level);
const orc_frame* orc_vo_ref_frame(const orc_vo*);
const orc_estimator* orc_vo_estimator(const orc_vo*);
int  orc_vo_trajectory(const orc_vo*, float* poses /* 16 per pose */, int max_poses);
/* last Result::pointCloud (vo.cc:260-281): xyzw, weight, gray per point */
int  orc_vo_point_cloud(const orc_vo*, float* xyzw, float* weights, uint8_t* gray, int max_points);
const char* orc_last_error(void);

/* ---- the disparity producer in front of the path: OpenCV's StereoBM as utils/stereo_algorithm.cc:67-111
 *      configures and calls it; restated in stereo_oracle.cc, pinned bit-exact against cv2 4.13 ---- */
typedef struct {
  int numberOfDisparities;   /* multiple of 16 */
  int SADWindowSize;         /* odd, >= 5 */
  int minDisparity;          /* <= 0 */
  int preFilterCap;          /* 1..63 (XSOBEL pre-filter) */
  int textureThreshold;
  int uniquenessRatio;
} orc_stereo_params;
void orc_stereo_prefilter_xsobel(const uint8_t* src, int rows, int cols, int cap, uint8_t* dst);
/* disp16: CV_16S fixed point (4 fractional bits), disp_f32 (may be NULL): disp16 / 16 as StereoAlgorithm::run returns it */
int orc_stereo_bm(const uint8_t* left, const uint8_t* right, int rows, int cols, const orc_stereo_params* p, int16_t* disp16, float* disp_f32);

#ifdef __cplusplus
}
#endif
#endif
