// vo_b200_seam.cc -- the reference-side binding of the B200 engine (Option A).
//
// Implements bpvo::VisualOdometryFrame and bpvo::VisualOdometryPoseEstimator -- the two classes the reference's bpvo/vo.cc
// drives -- on the C ABI of include/bpvo_b200.h.  Compile the reference's bpvo/vo.cc UNCHANGED
// against the two headers under integration/bpvo/ and link this file + libbpvo_b200.so: the key-frame state machine of vo.cc
// stays on the host, everything below setData / setTemplate / estimatePose runs on the GPU.
//
//   reference call (bpvo/vo.cc)                     C-ABI entry point
//   VisualOdometryFrame::setData        :131        bpvo_b200_frame_set_data
//   VisualOdometryFrame::setTemplate    :136,169,177 bpvo_b200_frame_set_template
//   getTemplateDataAtLevel(l)->points() :234,245,263 bpvo_b200_frame_num_points / bpvo_b200_frame_get_points
//   VisualOdometryPoseEstimator::estimatePose :144,184  bpvo_b200_estimate_pose
//   getFractionOfGoodPoints             :216        bpvo_b200_fraction_good
//   getWeights                          :264        bpvo_b200_get_weights (the first N entries = channel 0, all vo.cc reads)
//
// Context sharing: the reference constructs the pose estimator from the parameters alone and the frames from (K, b, params),
// none of them knows the image size (vo.cc:94-110).  The estimator therefore opens a Group that the frames constructed after
// it on the same thread join; the Group creates the engine context on the first setData(), when the size is known.
#include <bpvo/vo_frame.h>
#include <bpvo/vo_pose_estimator.h>
#include <bpvo/utils.h>
#include <bpvo/opencv.h>

#include <cstring>

#include "bpvo_b200.h"

namespace bpvo {
namespace b200 {

struct Group {
  bpvo_b200_ctx* ctx = nullptr;
  int rows = 0, cols = 0;
  ~Group() { if (ctx) bpvo_b200_destroy(ctx); }
};

static thread_local std::weak_ptr<Group> g_open_group;

static void check(int rc) { if (rc != BPVO_B200_OK) THROW_ERROR(bpvo_b200_last_error()); }

static bpvo_b200_params to_c_params(const AlgorithmParameters& p) {
  bpvo_b200_params q;
  bpvo_b200_default_params(&q);
  q.numPyramidLevels = p.numPyramidLevels; q.minImageDimensionForPyramid = p.minImageDimensionForPyramid;
  q.sigmaPriorToCensusTransform = p.sigmaPriorToCensusTransform; q.sigmaBitPlanes = p.sigmaBitPlanes;
  q.maxIterations = p.maxIterations; q.parameterTolerance = p.parameterTolerance; q.functionTolerance = p.functionTolerance;
  q.gradientTolerance = p.gradientTolerance; q.relaxTolerancesForCoarseLevels = p.relaxTolerancesForCoarseLevels ? 1 : 0;
  q.gradientEstimation = (int) p.gradientEstimation; q.interp = (int) p.interp;
  q.lossFunction = (int) p.lossFunction; q.descriptor = (int) p.descriptor; q.verbosity = (int) p.verbosity;
  q.minTranslationMagToKeyFrame = p.minTranslationMagToKeyFrame; q.minRotationMagToKeyFrame = p.minRotationMagToKeyFrame;
  q.maxFractionOfGoodPointsToKeyFrame = p.maxFractionOfGoodPointsToKeyFrame; q.goodPointThreshold = p.goodPointThreshold;
  q.minNumPixelsForNonMaximaSuppression = p.minNumPixelsForNonMaximaSuppression; q.nonMaxSuppRadius = p.nonMaxSuppRadius;
  q.minNumPixelsToWork = p.minNumPixelsToWork; q.minSaliency = p.minSaliency;
  q.minValidDisparity = p.minValidDisparity; q.maxValidDisparity = p.maxValidDisparity;
  q.maxTestLevel = p.maxTestLevel; q.withNormalization = p.withNormalization ? 1 : 0;
  q.dfSigma1 = p.dfSigma1; q.dfSigma2 = p.dfSigma2;
  return q;
}

}  // namespace b200

// ---- TemplateData view -------------------------------------------------------------------------------------------------
auto TemplateData::points() const -> const PointVector&
{
  if (!_fresh) {
    int n = 0;
    b200::check(bpvo_b200_frame_num_points(_frame, _level, &n));
    _points.resize(n);
    if (n > 0) b200::check(bpvo_b200_frame_get_points(_frame, _level, reinterpret_cast<float*>(_points.data())));   // Point = 4 packed floats
    _fresh = true;
  }
  return _points;
}

// ---- VisualOdometryFrame (vo_frame.cc:13-93) -------------------------------------------------------------------------------
VisualOdometryFrame::VisualOdometryFrame(const Matrix33& K, float b, const AlgorithmParameters& p)
    : _K(K), _b(b), _params(p), _has_data(false), _has_template(false), _group(b200::g_open_group.lock()), _h(nullptr)
{
  if (!_group) _group = std::make_shared<b200::Group>();       // a frame used without a pose estimator
}

VisualOdometryFrame::~VisualOdometryFrame() { if (_h) bpvo_b200_frame_destroy(_h); }

void VisualOdometryFrame::sync_flags() { if (_h && !_has_data) bpvo_b200_frame_clear(_h); for (auto& t : _tdata) t._fresh = false; }

void VisualOdometryFrame::setData(const cv::Mat& image, const cv::Mat& disparity)
{
  b200::Group& g = *_group;
  if (!g.ctx) {                                                // first frame of this VisualOdometry: the image size is known now
    const bpvo_b200_params q = b200::to_c_params(_params);
    g.rows = image.rows; g.cols = image.cols;
    b200::check(bpvo_b200_create(&g.ctx, _K.data(), _b, image.rows, image.cols, &q));       // Matrix33::data(): column-major
  }
  THROW_ERROR_IF(image.rows != g.rows || image.cols != g.cols, "image size changed");
  if (!_h) {
    b200::check(bpvo_b200_frame_create(g.ctx, &_h));
    const int L = bpvo_b200_frame_num_levels(_h);
    _tdata.resize(L);
    Matrix33 Kl(_K);
    for (int l = 0; l < L; ++l) {                              // K_l = K / 2^l with K(2,2) = 1 (vo_frame.cc:24-28)
      if (l > 0) { Kl *= 0.5f; Kl(2, 2) = 1.0f; }
      _tdata[l]._frame = _h; _tdata[l]._level = l; _tdata[l]._warp.K = Kl;
    }
  }
  _image = make_unique<cv::Mat>(image.clone());                // vo.cc colours the point cloud from the raw image (vo_frame.cc:50)
  b200::check(bpvo_b200_frame_set_data(_h, image.ptr<uint8_t>(), disparity.ptr<float>()));
  _has_data = true;
}

void VisualOdometryFrame::setTemplate()
{
  THROW_ERROR_IF(!_has_data || !_h, "no data in frame");      // vo_frame.cc:63
  b200::check(bpvo_b200_frame_set_template(_h));
  for (auto& t : _tdata) t._fresh = false;
  _has_template = true;
}

const TemplateData* VisualOdometryFrame::getTemplateDataAtLevel(size_t l) const
{
  THROW_ERROR_IF(l >= _tdata.size(), "no template data at this level");
  return &_tdata[l];
}

int VisualOdometryFrame::numLevels() const { return _h ? bpvo_b200_frame_num_levels(_h) : _params.numPyramidLevels; }

const cv::Mat* VisualOdometryFrame::imagePointer() const { return _image.get(); }

// ---- VisualOdometryPoseEstimator (vo_pose_estimator.cc:55-107) -----------------------------------------------------------
VisualOdometryPoseEstimator::VisualOdometryPoseEstimator(const AlgorithmParameters& p)
    : _params(p), _group(std::make_shared<b200::Group>()), _weights_fresh(false)
{
  b200::g_open_group = _group;                                 // the frames constructed next on this thread join it (vo.cc:98-109)
}

VisualOdometryPoseEstimator::~VisualOdometryPoseEstimator() {}

std::vector<OptimizerStatistics> VisualOdometryPoseEstimator::
estimatePose(const VisualOdometryFrame* ref_frame, const VisualOdometryFrame* cur_frame, const Matrix44& T_init, Matrix44& T_est)
{
  THROW_ERROR_IF(!_group->ctx || !ref_frame->handle() || !cur_frame->handle(), "you should call setData before calling computeResiduals");
  bpvo_b200_stats st[BPVO_B200_MAX_LEVELS];
  b200::check(bpvo_b200_estimate_pose(_group->ctx, ref_frame->handle(), cur_frame->handle(), T_init.data(), T_est.data(), st, nullptr));
  const int L = ref_frame->numLevels();
  std::vector<OptimizerStatistics> ret(L);
  for (int l = 0; l < L; ++l) {
    ret[l].numIterations = st[l].numIterations; ret[l].finalError = st[l].finalError;
    ret[l].firstOrderOptimality = st[l].firstOrderOptimality; ret[l].status = static_cast<PoseEstimationStatus>(st[l].status);
  }
  _weights_fresh = false;
  return ret;
}

float VisualOdometryPoseEstimator::getFractionOfGoodPoints(float thresh) const
{
  float f = 0.0f;
  b200::check(bpvo_b200_fraction_good(_group->ctx, thresh, &f));
  return f;
}

const WeightsVector& VisualOdometryPoseEstimator::getWeights() const
{
  if (!_weights_fresh) {
    size_t n = 0;
    b200::check(bpvo_b200_get_weights(_group->ctx, nullptr, &n));
    _weights.resize(n);
    size_t cap = n;
    if (n > 0) b200::check(bpvo_b200_get_weights(_group->ctx, _weights.data(), &cap));
    _weights_fresh = true;
  }
  return _weights;
}

}  // namespace bpvo
