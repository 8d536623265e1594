// ref_shim.cc -- C wrappers around the REAL reference code (compiled from where it lies under
// /root/reference/bpvo, against the header stand-ins in oracle/refstub/) for the leaf translation units
// that build without Eigen/OpenCV proper: mestimator.cc, census.cc, imgproc.cc, linear_system_builder.cc,
// utils.cc (+ the header-only IsLocalMax and median).  Output: oracle/_ref/libbpvo_ref.so.
// Purpose: pin the oracle's restatement of exactly these quirk-laden SIMD pieces against the reference
// itself.  TEST INFRASTRUCTURE ONLY; no reference source is copied into this repository.
#include <bpvo/census.h>
#include <bpvo/imgproc.h>
#include <bpvo/linear_system_builder.h>
#include <bpvo/mestimator.h>
#include <bpvo/types.h>
#include <bpvo/utils.h>

#include <cstring>

using namespace bpvo;

extern "C" {

// bpvo::census (census.cc:59-91)
void ref_census(const uint8_t* src, int rows, int cols, uint8_t* dst) {
  cv::Mat I(rows, cols, CV_8UC1, (void*) src);
  cv::Mat C = census(I, -1.0f);
  memcpy(dst, C.ptr<uint8_t>(), (size_t) rows * cols);
}

// bpvo::census with the pre-census blur (census.cc:63-65): sigma > 0 runs cv::GaussianBlur(3x3) of the stand-in OpenCV
void ref_census_sigma(const uint8_t* src, int rows, int cols, float sigma, uint8_t* dst) {
  cv::Mat I(rows, cols, CV_8UC1, (void*) src);
  cv::Mat C = census(I, sigma);
  memcpy(dst, C.ptr<uint8_t>(), (size_t) rows * cols);
}

// DenseDescriptor::computeSaliencyMap (dense_descriptor.cc:92-100) driving the real
// gradientAbsoluteMagnitude / gradientAbsoluteMagnitudeAcc (imgproc.cc:45-142)
void ref_saliency(const float* planes, int channels, int rows, int cols, float* dst) {
  cv::Mat_<float> d;
  d.create(rows, cols);
  cv::Mat_<float> c0(rows, cols, const_cast<float*>(planes));
  gradientAbsoluteMagnitude(c0, d);
  for (int i = 1; i < channels; ++i) {
    cv::Mat_<float> ci(rows, cols, const_cast<float*>(planes) + (size_t) i * rows * cols);
    gradientAbsoluteMagnitudeAcc(ci, d.ptr<float>());
  }
  memcpy(dst, d.ptr<float>(), (size_t) rows * cols * sizeof(float));
}

// IsLocalMax<float> (imgproc.h:93-165) over the candidate window used by TemplateData::setData
void ref_local_max(const float* S, int rows, int cols, int radius, int border, uint8_t* out) {
  IsLocalMax<float> f(S, cols, radius);
  memset(out, 0, (size_t) rows * cols);
  for (int y = border; y < rows - border - 1; ++y)
    for (int x = border; x < cols - border - 1; ++x) out[(size_t) y * cols + x] = f(y, x) ? 1 : 0;
}

// median (utils.h:224-252)
float ref_median(const float* buf, size_t n) {
  std::vector<float> v(buf, buf + n);
  return median(v);
}

// MEstimator::ComputeWeights (mestimator.cc:390-415)
void ref_compute_weights(int loss, const float* r, const uint16_t* valid, size_t n, float sigma, float* w) {
  ResidualsVector R(r, r + n); ValidVector V(valid, valid + n); WeightsVector W;
  MEstimator::ComputeWeights((LossFunctionType) loss, R, V, sigma, W);
  memcpy(w, W.data(), n * sizeof(float));
}

// AutoScaleEstimator (mestimator.cc:416-490)
void* ref_scale_create() { return new AutoScaleEstimator(); }
void ref_scale_destroy(void* h) { delete (AutoScaleEstimator*) h; }
void ref_scale_reset(void* h) { ((AutoScaleEstimator*) h)->reset(); }
float ref_scale_estimate(void* h, const float* r, const uint16_t* valid, size_t n) {
  ResidualsVector R(r, r + n); ValidVector V(valid, valid + n);
  return ((AutoScaleEstimator*) h)->estimateScale(R, V);
}

// LinearSystemBuilder::Run (linear_system_builder.cc:334-350); J is n x 6 row-major, a zero Jacobian is appended
// as TemplateData::setData does (template_data.cc:139-141); H comes back column-major
float ref_linear_system(const float* J, const float* r, const float* w, const uint16_t* valid, size_t n, float* H, float* G) {
  LinearSystemBuilder::JacobianVector Jv(n + 1);
  for (size_t i = 0; i < n; ++i) memcpy(Jv[i].data(), J + 6 * i, 6 * sizeof(float));
  Jv[n].setZero();
  ResidualsVector R(r, r + n), W(w, w + n); ValidVector V(valid, valid + n);
  LinearSystemBuilder::Hessian Hm; LinearSystemBuilder::Gradient Gm;
  float f = LinearSystemBuilder::Run(Jv, R, W, V, &Hm, &Gm);
  memcpy(H, Hm.data(), 36 * sizeof(float)); memcpy(G, Gm.data(), 6 * sizeof(float));
  return f;
}

}  // extern "C"

// =====================================================================================================================
// The whole hot path of the REAL reference (vo.cc, vo_frame.cc, vo_pose_estimator.cc, pose_estimator_{base,gn}.h,
// template_data.cc, rigid_body_warp.{h,cc}, warps.cc, photo_error.cc, dense_descriptor*.cc, bitplanes/intensity
// descriptors, image_pyramid.cc, trajectory.cc, point_cloud.cc, types.cc, ...) compiled from where it lies, against the
// stand-in Eigen/OpenCV headers of oracle/refstub.  Third-party arithmetic (matrix products, LDLT, 4x4 inverse, pyrDown,
// GaussianBlur) is therefore the stand-ins', everything else -- including the SSE rcp Jacobians, the fp64 photo error,
// the selection glue, the GN loop with its quirks and the key-frame state machine -- is the reference's own code.
// =====================================================================================================================
#include <bpvo/dense_descriptor.h>
#include <bpvo/parallel.h>
#include <bpvo/point_cloud.h>
#include <bpvo/pose_estimator_gn.h>
#include <bpvo/template_data.h>
#include <bpvo/trajectory.h>
#include <bpvo/vo.h>
#include <bpvo/vo_frame.h>
#include <bpvo/vo_pose_estimator.h>

#include "bpvo_oracle.h"     // POD layouts shared with the oracle's C API (orc_params, orc_stats, orc_result)

namespace {

thread_local std::string g_ref_err;

AlgorithmParameters to_params(const orc_params* q) {
  AlgorithmParameters p;
  p.numPyramidLevels = q->numPyramidLevels; p.minImageDimensionForPyramid = q->minImageDimensionForPyramid;
  p.sigmaPriorToCensusTransform = q->sigmaPriorToCensusTransform; p.sigmaBitPlanes = q->sigmaBitPlanes;
  p.maxIterations = q->maxIterations; p.parameterTolerance = q->parameterTolerance; p.functionTolerance = q->functionTolerance;
  p.gradientTolerance = q->gradientTolerance; p.relaxTolerancesForCoarseLevels = q->relaxTolerancesForCoarseLevels != 0;
  p.gradientEstimation = (GradientEstimationType) q->gradientEstimation; p.interp = (InterpolationType) q->interp;
  p.lossFunction = (LossFunctionType) q->lossFunction; p.descriptor = (DescriptorType) q->descriptor;
  p.verbosity = VerbosityType::kSilent;
  p.minTranslationMagToKeyFrame = q->minTranslationMagToKeyFrame; p.minRotationMagToKeyFrame = q->minRotationMagToKeyFrame;
  p.maxFractionOfGoodPointsToKeyFrame = q->maxFractionOfGoodPointsToKeyFrame; p.goodPointThreshold = q->goodPointThreshold;
  p.minNumPixelsForNonMaximaSuppression = q->minNumPixelsForNonMaximaSuppression; p.nonMaxSuppRadius = q->nonMaxSuppRadius;
  p.minNumPixelsToWork = q->minNumPixelsToWork; p.minSaliency = q->minSaliency;
  p.minValidDisparity = q->minValidDisparity; p.maxValidDisparity = q->maxValidDisparity;
  p.maxTestLevel = q->maxTestLevel; p.withNormalization = q->withNormalization != 0;
  p.dfSigma1 = q->dfSigma1; p.dfSigma2 = q->dfSigma2;
  return p;
}

Matrix33 to_K(const float K[9]) { Matrix33 m; memcpy(m.data(), K, 9 * sizeof(float)); return m; }

struct GnAccess : public PoseEstimatorGN<TemplateData> {      // exposes the protected residual / valid vectors
  using PoseEstimatorGN<TemplateData>::residuals;
  using PoseEstimatorGN<TemplateData>::valid;
  using PoseEstimatorGN<TemplateData>::weights;
  void do_reset() { this->reset(); }
};

struct RefFrame { AlgorithmParameters p; int rows, cols; std::unique_ptr<VisualOdometryFrame> f; };
struct RefEst { AlgorithmParameters p; GnAccess gn; float sigma = 1.0f; };
struct RefVo { std::unique_ptr<VisualOdometry> vo; int rows, cols, levels; std::unique_ptr<PointCloud> cloud; };

}  // namespace

#define REF_TRY try {
#define REF_CATCH(rv) } catch (const std::exception& e) { g_ref_err = e.what(); return rv; }

extern "C" {

const char* ref_last_error() { return g_ref_err.c_str(); }
// bpvo::setNumThreads (bpvo/parallel.cc:112-126): threads of the reference's OpenMP parallel_for
void ref_set_num_threads(int n) { bpvo::setNumThreads(n); }
int ref_get_num_threads() { return bpvo::getNumThreads(); }

// ---- VisualOdometryFrame (vo_frame.cc) -----------------------------------------------------------------------------
void* ref_frame_create(const float K[9], float baseline, int rows, int cols, const orc_params* q) {
  REF_TRY
  std::unique_ptr<RefFrame> h(new RefFrame);
  h->p = to_params(q); h->rows = rows; h->cols = cols;
  h->f.reset(new VisualOdometryFrame(to_K(K), baseline, h->p));
  return h.release();
  REF_CATCH(nullptr)
}
void ref_frame_destroy(void* h) { delete (RefFrame*) h; }
int ref_frame_set_data(void* hh, const uint8_t* image, const float* disparity) {
  REF_TRY
  RefFrame* h = (RefFrame*) hh;
  cv::Mat I(h->rows, h->cols, CV_8UC1, (void*) image), D(h->rows, h->cols, CV_32FC1, (void*) disparity);
  h->f->setData(I, D);
  return 0;
  REF_CATCH(-1)
}
int ref_frame_set_template(void* hh) { REF_TRY ((RefFrame*) hh)->f->setTemplate(); return 0; REF_CATCH(-1) }
int ref_frame_num_points(void* hh, int l) { return ((RefFrame*) hh)->f->getTemplateDataAtLevel(l)->numPoints(); }
void ref_frame_level_size(void* hh, int l, int* rows, int* cols) {
  const DenseDescriptor* d = ((RefFrame*) hh)->f->getDenseDescriptorAtLevel(l); *rows = d->rows(); *cols = d->cols();
}
int ref_frame_descriptor(void* hh, int l, float* planes) {
  const DenseDescriptor* d = ((RefFrame*) hh)->f->getDenseDescriptorAtLevel(l);
  const size_t n = (size_t) d->rows() * d->cols();
  for (int c = 0; c < d->numChannels(); ++c) memcpy(planes + c * n, d->getChannel(c).ptr<float>(), n * sizeof(float));
  return d->numChannels();
}
void ref_frame_points(void* hh, int l, float* xyzw) {
  const TemplateData* t = ((RefFrame*) hh)->f->getTemplateDataAtLevel(l);
  if (t->numPoints()) memcpy(xyzw, t->points()[0].data(), (size_t) t->numPoints() * 16);
}
void ref_frame_pixels(void* hh, int l, float* out) {
  const TemplateData* t = ((RefFrame*) hh)->f->getTemplateDataAtLevel(l);
  memcpy(out, t->pixels().data(), t->pixels().size() * sizeof(float));
}
void ref_frame_jacobians(void* hh, int l, float* out) {     // C*N rows of 6 (the trailing zero Jacobian is dropped)
  const TemplateData* t = ((RefFrame*) hh)->f->getTemplateDataAtLevel(l);
  const size_t n = t->jacobians().size() - 1;
  for (size_t i = 0; i < n; ++i) memcpy(out + 6 * i, t->jacobians()[i].data(), 6 * sizeof(float));
}
int ref_frame_num_pixels(void* hh, int l) { return ((RefFrame*) hh)->f->getTemplateDataAtLevel(l)->numPixels(); }

// ---- PoseEstimatorGN<TemplateData> (pose_estimator_gn.h / pose_estimator_base.h) ---------------------------------------
void* ref_estimator_create(const orc_params* q) {
  REF_TRY
  std::unique_ptr<RefEst> e(new RefEst);
  e->p = to_params(q);
  e->gn.setParameters(PoseEstimatorParameters(e->p));
  return e.release();
  REF_CATCH(nullptr)
}
void ref_estimator_destroy(void* e) { delete (RefEst*) e; }
// one PoseEstimatorGN::linearize at pose T (column-major); reset != 0 calls PoseEstimatorBase::reset() first
float ref_linearize(void* ee, void* ref, void* cur, int level, const float T[16], int reset, float H[36], float G[6]) {
  REF_TRY
  RefEst* e = (RefEst*) ee;
  if (reset) e->gn.do_reset();
  typename GnAccess::PoseEstimatorData data;
  memcpy(data.T.data(), T, 16 * sizeof(float));
  const TemplateData* td = ((RefFrame*) ref)->f->getTemplateDataAtLevel(level);
  const DenseDescriptor* dd = ((RefFrame*) cur)->f->getDenseDescriptorAtLevel(level);
  const float f = e->gn.linearize(td, dd, data);
  memcpy(H, data.H.data(), 36 * sizeof(float)); memcpy(G, data.G.data(), 6 * sizeof(float));
  return f;
  REF_CATCH(-1.0f)
}
size_t ref_estimator_num_residuals(void* ee) { return ((RefEst*) ee)->gn.residuals().size(); }
void ref_estimator_vectors(void* ee, float* r, float* w, uint16_t* v) {
  RefEst* e = (RefEst*) ee;
  const size_t n = e->gn.residuals().size();
  if (r) memcpy(r, e->gn.residuals().data(), n * sizeof(float));
  if (w) memcpy(w, e->gn.weights().data(), n * sizeof(float));
  if (v) memcpy(v, e->gn.valid().data(), n * sizeof(uint16_t));
}
// PoseEstimatorBase::run at one level (the reference's own GN loop incl. solve(), testConvergence(), Q1/Q2)
int ref_run_level(void* ee, void* ref, void* cur, int level, float T[16], orc_stats* st) {
  REF_TRY
  RefEst* e = (RefEst*) ee;
  Matrix44 Tm; memcpy(Tm.data(), T, 16 * sizeof(float));
  const TemplateData* td = ((RefFrame*) ref)->f->getTemplateDataAtLevel(level);
  const DenseDescriptor* dd = ((RefFrame*) cur)->f->getDenseDescriptorAtLevel(level);
  OptimizerStatistics s = e->gn.run(td, dd, Tm);
  memcpy(T, Tm.data(), 16 * sizeof(float));
  st->numIterations = s.numIterations; st->finalError = s.finalError; st->firstOrderOptimality = s.firstOrderOptimality; st->status = (int) s.status;
  return 0;
  REF_CATCH(-1)
}

// ---- VisualOdometry (vo.cc) ------------------------------------------------------------------------------------------
void* ref_vo_create(const float K[9], float baseline, int rows, int cols, const orc_params* q) {
  REF_TRY
  std::unique_ptr<RefVo> h(new RefVo);
  AlgorithmParameters p = to_params(q);
  h->rows = rows; h->cols = cols;
  h->vo.reset(new VisualOdometry(to_K(K), baseline, ImageSize(rows, cols), p));
  return h.release();
  REF_CATCH(nullptr)
}
void ref_vo_destroy(void* h) { delete (RefVo*) h; }
int ref_vo_add_frame(void* hh, const uint8_t* image, const float* disparity, orc_result* out) {
  REF_TRY
  RefVo* h = (RefVo*) hh;
  Result r = h->vo->addFrame(image, disparity);
  memset(out, 0, sizeof(*out));
  memcpy(out->pose, r.pose.data(), 16 * sizeof(float));
  out->isKeyFrame = r.isKeyFrame ? 1 : 0; out->keyFramingReason = (int) r.keyFramingReason;
  out->numLevels = (int) r.optimizerStatistics.size();
  for (int i = 0; i < out->numLevels && i < 16; ++i) {
    out->stats[i].numIterations = r.optimizerStatistics[i].numIterations; out->stats[i].finalError = r.optimizerStatistics[i].finalError;
    out->stats[i].firstOrderOptimality = r.optimizerStatistics[i].firstOrderOptimality; out->stats[i].status = (int) r.optimizerStatistics[i].status;
  }
  out->numPointCloud = r.pointCloud ? (int) r.pointCloud->size() : 0;
  h->cloud = std::move(r.pointCloud);
  return 0;
  REF_CATCH(-1)
}
int ref_vo_num_points_at_level(void* hh, int level) { return ((RefVo*) hh)->vo->numPointsAtLevel(level); }
int ref_vo_trajectory(void* hh, float* poses, int max_poses) {
  const Trajectory& t = ((RefVo*) hh)->vo->trajectory();
  const int n = (int) t.size();
  for (int i = 0; i < n && i < max_poses; ++i) memcpy(poses + 16 * (size_t) i, t[i].data(), 16 * sizeof(float));
  return n;
}
int ref_vo_point_cloud(void* hh, float* xyzw, float* weights, uint8_t* gray, int max_points) {
  RefVo* h = (RefVo*) hh;
  const int n = h->cloud ? (int) h->cloud->size() : 0;
  for (int i = 0; i < n && i < max_points; ++i) {
    const PointWithInfo& p = (*h->cloud)[i];
    memcpy(xyzw + 4 * (size_t) i, p.xyzw().data(), 16); weights[i] = p.weight(); gray[i] = p.rgba()[0];
  }
  return n;
}

}  // extern "C"
