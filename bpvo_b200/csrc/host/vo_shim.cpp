// host/vo_shim.cpp -- bpvo/vo.cc (VisualOdometry::Impl, :94-281) restated on top of the seam-level C ABI,
// plus the VisualOdometry-level C entry points (bpvo_b200_vo_*) that a MEX / ctypes / cgo binding uses.
// Uses ONLY functions declared in include/bpvo_b200.h -- it is the drop-in demonstration.
#include "vo.h"

#include <math.h>
#include <string.h>
#include <algorithm>
#include <utility>

namespace bpvo_b200 {

Matrix44 Matrix44::Identity() { Matrix44 r; memset(r.m, 0, sizeof(r.m)); r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f; return r; }

static Matrix44 mul(const Matrix44& a, const Matrix44& b) {
  Matrix44 r;
  for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) {
    float s = a(i, 0) * b(0, j);
    for (int k = 1; k < 4; ++k) s += a(i, k) * b(k, j);
    r(i, j) = s;
  }
  return r;
}

// Matrix4f::inverse() (general adjugate)
static Matrix44 inverse(const Matrix44& a) {
  const float* m = a.m; float inv[16];
  inv[0] = m[5]*m[10]*m[15] - m[5]*m[11]*m[14] - m[9]*m[6]*m[15] + m[9]*m[7]*m[14] + m[13]*m[6]*m[11] - m[13]*m[7]*m[10];
  inv[4] = -m[4]*m[10]*m[15] + m[4]*m[11]*m[14] + m[8]*m[6]*m[15] - m[8]*m[7]*m[14] - m[12]*m[6]*m[11] + m[12]*m[7]*m[10];
  inv[8] = m[4]*m[9]*m[15] - m[4]*m[11]*m[13] - m[8]*m[5]*m[15] + m[8]*m[7]*m[13] + m[12]*m[5]*m[11] - m[12]*m[7]*m[9];
  inv[12] = -m[4]*m[9]*m[14] + m[4]*m[10]*m[13] + m[8]*m[5]*m[14] - m[8]*m[6]*m[13] - m[12]*m[5]*m[10] + m[12]*m[6]*m[9];
  inv[1] = -m[1]*m[10]*m[15] + m[1]*m[11]*m[14] + m[9]*m[2]*m[15] - m[9]*m[3]*m[14] - m[13]*m[2]*m[11] + m[13]*m[3]*m[10];
  inv[5] = m[0]*m[10]*m[15] - m[0]*m[11]*m[14] - m[8]*m[2]*m[15] + m[8]*m[3]*m[14] + m[12]*m[2]*m[11] - m[12]*m[3]*m[10];
  inv[9] = -m[0]*m[9]*m[15] + m[0]*m[11]*m[13] + m[8]*m[1]*m[15] - m[8]*m[3]*m[13] - m[12]*m[1]*m[11] + m[12]*m[3]*m[9];
  inv[13] = m[0]*m[9]*m[14] - m[0]*m[10]*m[13] - m[8]*m[1]*m[14] + m[8]*m[2]*m[13] + m[12]*m[1]*m[10] - m[12]*m[2]*m[9];
  inv[2] = m[1]*m[6]*m[15] - m[1]*m[7]*m[14] - m[5]*m[2]*m[15] + m[5]*m[3]*m[14] + m[13]*m[2]*m[7] - m[13]*m[3]*m[6];
  inv[6] = -m[0]*m[6]*m[15] + m[0]*m[7]*m[14] + m[4]*m[2]*m[15] - m[4]*m[3]*m[14] - m[12]*m[2]*m[7] + m[12]*m[3]*m[6];
  inv[10] = m[0]*m[5]*m[15] - m[0]*m[7]*m[13] - m[4]*m[1]*m[15] + m[4]*m[3]*m[13] + m[12]*m[1]*m[7] - m[12]*m[3]*m[5];
  inv[14] = -m[0]*m[5]*m[14] + m[0]*m[6]*m[13] + m[4]*m[1]*m[14] - m[4]*m[2]*m[13] - m[12]*m[1]*m[6] + m[12]*m[2]*m[5];
  inv[3] = -m[1]*m[6]*m[11] + m[1]*m[7]*m[10] + m[5]*m[2]*m[11] - m[5]*m[3]*m[10] - m[9]*m[2]*m[7] + m[9]*m[3]*m[6];
  inv[7] = m[0]*m[6]*m[11] - m[0]*m[7]*m[10] - m[4]*m[2]*m[11] + m[4]*m[3]*m[10] + m[8]*m[2]*m[7] - m[8]*m[3]*m[6];
  inv[11] = -m[0]*m[5]*m[11] + m[0]*m[7]*m[9] + m[4]*m[1]*m[11] - m[4]*m[3]*m[9] - m[8]*m[1]*m[7] + m[8]*m[3]*m[5];
  inv[15] = m[0]*m[5]*m[10] - m[0]*m[6]*m[9] - m[4]*m[1]*m[10] + m[4]*m[2]*m[9] + m[8]*m[1]*m[6] - m[8]*m[2]*m[5];
  const float det = m[0]*inv[0] + m[1]*inv[4] + m[2]*inv[8] + m[3]*inv[12];
  const float idet = 1.0f / det;
  Matrix44 r; for (int i = 0; i < 16; ++i) r.m[i] = inv[i] * idet;
  return r;
}

Result::Result() : pose(Matrix44::Identity()) {
  memset(covariance, 0, sizeof(covariance));
  for (int i = 0; i < 6; ++i) covariance[i * 6 + i] = 1.0f;
}

// Trajectory::push_back with its InvertPose (bpvo/trajectory.cc:30-50); the translation of the "inverse" is
// -R*t (sic: the reference transposes twice), kept as is
void Trajectory::push_back(const Matrix44& T) {
  Matrix44 inv; memset(inv.m, 0, sizeof(inv.m));
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) inv(i, j) = T(j, i);
  for (int i = 0; i < 3; ++i) { float s = T(i, 0) * T(0, 3); s += T(i, 1) * T(1, 3); s += T(i, 2) * T(2, 3); inv(i, 3) = -s; }
  inv(3, 3) = 1.0f;
  if (!_poses.empty()) _poses.push_back(mul(_poses.back(), inv)); else _poses.push_back(inv);
}

static void check(int rc) { if (rc != BPVO_B200_OK) throw Error(bpvo_b200_last_error()); }

class VisualOdometry::Impl {
 public:
  Impl(const Matrix33& K, float b, ImageSize s, const AlgorithmParameters& p) : _params(p), _image_size(s), _K(K) {
    if (_params.numPyramidLevels <= 0) {                                           // vo.cc:101-104
      _params.numPyramidLevels = 1 + (int) std::round(std::log2(std::min(s.rows, s.cols) / (double) p.minImageDimensionForPyramid));
    }
    check(bpvo_b200_create(&_ctx, K.data(), b, s.rows, s.cols, &_params));          // VisualOdometryPoseEstimator (vo.cc:98)
    try {
      check(bpvo_b200_frame_create(_ctx, &_ref_frame));                             // vo.cc:107-109
      check(bpvo_b200_frame_create(_ctx, &_cur_frame));
      check(bpvo_b200_frame_create(_ctx, &_prev_frame));
    } catch (...) { destroy(); throw; }
    _T_kf = Matrix44::Identity();
  }
  ~Impl() { destroy(); }

  // VisualOdometry::Impl::addFrame (vo.cc:125-197)
  Result addFrame(const uint8_t* I_ptr, const float* D_ptr) {
    Result ret;
    check(bpvo_b200_frame_set_data(_cur_frame, I_ptr, D_ptr));                      // vo.cc:131
    const int n_levels = bpvo_b200_frame_num_levels(_cur_frame);
    if (!bpvo_b200_frame_has_template(_ref_frame)) {                               // vo.cc:133-139
      std::swap(_ref_frame, _cur_frame);
      check(bpvo_b200_frame_set_template(_ref_frame));
      _trajectory.push_back(_T_kf);
      ret.optimizerStatistics.resize(n_levels);                                    // FirstFrameResult (vo.cc:112-123)
      for (auto& s : ret.optimizerStatistics) { s.numIterations = 0; s.finalError = -1.0f; s.firstOrderOptimality = -1.0f; s.status = BPVO_B200_SOLVER_ERROR; }
      ret.isKeyFrame = true; ret.keyFramingReason = BPVO_B200_KF_FIRST_FRAME;
      _points_dirty = true;
      check(bpvo_b200_synchronize(_ctx));    // the borrowed (pinned) input buffers are free again when addFrame returns
      return ret;
    }
    Matrix44 T_est;
    ret.optimizerStatistics.resize(n_levels);
    int evals = 0;
    check(bpvo_b200_estimate_pose(_ctx, _ref_frame, _cur_frame, _T_kf.data(), T_est.data(), ret.optimizerStatistics.data(), &evals));   // vo.cc:144
    ret.numFunEvals += evals;
    ret.keyFramingReason = shouldKeyFrame(T_est);
    ret.isKeyFrame = BPVO_B200_KF_NONE != ret.keyFramingReason;
    if (!ret.isKeyFrame) {                                                          // vo.cc:149-155
      std::swap(_prev_frame, _cur_frame);
      ret.pose = mul(T_est, inverse(_T_kf));
      _T_kf = T_est;
    } else {
      ret.pointCloud = getPointCloudFromRefFrame();                                 // vo.cc:159
      if (bpvo_b200_frame_empty(_prev_frame)) {                                     // vo.cc:161-173
        std::swap(_cur_frame, _ref_frame);
        check(bpvo_b200_frame_set_template(_ref_frame));
        ret.pose = mul(T_est, inverse(_T_kf));
        _T_kf = Matrix44::Identity();
      } else {                                                                      // vo.cc:175-187
        std::swap(_prev_frame, _ref_frame);
        check(bpvo_b200_frame_clear(_prev_frame));
        check(bpvo_b200_frame_set_template(_ref_frame));
        const Matrix44 T_init = Matrix44::Identity();
        check(bpvo_b200_estimate_pose(_ctx, _ref_frame, _cur_frame, T_init.data(), T_est.data(), ret.optimizerStatistics.data(), &evals));
        ret.numFunEvals += evals;
        ret.pose = T_est;
        _T_kf = T_est;
      }
      _points_dirty = true;
    }
    _trajectory.push_back(ret.pose);                                                // vo.cc:191
    if (ret.pointCloud) ret.pointCloud->pose = _trajectory.back();
    return ret;
  }

  // shouldKeyFrame (vo.cc:199-224); RotationMatrixToEulerAngles (math_utils.h:210-222); radians vs "degrees" (Q8) kept
  int shouldKeyFrame(const Matrix44& pose) const {
    const float t_norm = pose(0, 3) * pose(0, 3) + pose(1, 3) * pose(1, 3) + pose(2, 3) * pose(2, 3);
    if (t_norm > _params.minTranslationMagToKeyFrame * _params.minTranslationMagToKeyFrame) return BPVO_B200_KF_LARGE_TRANSLATION;
    const float eta = (float) (1.0 / (std::sqrt(pose(0, 0) * pose(0, 0) + pose(1, 0) * pose(1, 0))));
    const float rz = std::asin(eta * pose(1, 0)), ry = std::asin(-pose(2, 0)), rx = std::asin(eta * pose(2, 1));
    const float r_norm = rx * rx + ry * ry + rz * rz;
    if (r_norm > _params.minRotationMagToKeyFrame * _params.minRotationMagToKeyFrame) return BPVO_B200_KF_LARGE_ROTATION;
    float frac_good = 0.0f;
    check(bpvo_b200_fraction_good(_ctx, _params.goodPointThreshold, &frac_good));   // vo.cc:216
    if (frac_good < _params.maxFractionOfGoodPointsToKeyFrame) return BPVO_B200_KF_SMALL_FRAC_GOOD;
    return BPVO_B200_KF_NONE;
  }

  int numPointsAtLevel(int level) const {                                           // vo.cc:226-238
    if (level < 0) level = _params.maxTestLevel;
    int n = 0;
    if (_ref_frame) check(bpvo_b200_frame_num_points(_ref_frame, level, &n));
    return n;
  }

  const PointVector& pointsAtLevel(int level) const {                               // vo.cc:240-247
    if (!_ref_frame) throw Error("no reference frame has been set");
    if (level < 0) level = _params.maxTestLevel;
    if (_points_dirty || level != _points_level) {
      int n = 0; check(bpvo_b200_frame_num_points(_ref_frame, level, &n));
      _points.resize(n);
      if (n > 0) check(bpvo_b200_frame_get_points(_ref_frame, level, &_points[0].x));
      _points_dirty = false; _points_level = level;
    }
    return _points;
  }

  // getPointCloudFromRefFrame (vo.cc:249-281): weights[i], i < N, are channel 0 of the last linearize (Q7)
  std::unique_ptr<PointCloud> getPointCloudFromRefFrame() const {
    // assembled on the device (xyzw, grey level at K_l X in the full-resolution ref image, channel-0 weight): one download
    static_assert(sizeof(PointWithInfo) == sizeof(bpvo_b200_point_info), "PointWithInfo is the 32-byte device record");
    std::unique_ptr<PointCloud> ret(new PointCloud);
    ret->pose = Matrix44::Identity();
    int n = 0;
    check(bpvo_b200_point_cloud(_ctx, _ref_frame, nullptr, &n));
    ret->points.resize((size_t) n);
    if (n) { int cap = n; check(bpvo_b200_point_cloud(_ctx, _ref_frame, reinterpret_cast<bpvo_b200_point_info*>(ret->points.data()), &cap)); }
    return ret;
  }

  const Trajectory& trajectory() const { return _trajectory; }
  bpvo_b200_ctx* ctx() const { return _ctx; }
  const bpvo_b200_frame* refFrame() const { return _ref_frame; }

 private:
  void destroy() {
    if (_ref_frame) bpvo_b200_frame_destroy(_ref_frame);
    if (_cur_frame) bpvo_b200_frame_destroy(_cur_frame);
    if (_prev_frame) bpvo_b200_frame_destroy(_prev_frame);
    if (_ctx) bpvo_b200_destroy(_ctx);
    _ref_frame = _cur_frame = _prev_frame = nullptr; _ctx = nullptr;
  }
  AlgorithmParameters _params;
  ImageSize _image_size;
  Matrix33 _K;
  bpvo_b200_ctx* _ctx = nullptr;
  bpvo_b200_frame* _ref_frame = nullptr;
  bpvo_b200_frame* _cur_frame = nullptr;
  bpvo_b200_frame* _prev_frame = nullptr;
  Matrix44 _T_kf;
  Trajectory _trajectory;
  mutable PointVector _points; mutable bool _points_dirty = true; mutable int _points_level = -1;
};

VisualOdometry::VisualOdometry(const Matrix33& K, float baseline, ImageSize s, const AlgorithmParameters& p) : _impl(new Impl(K, baseline, s, p)) {}
VisualOdometry::~VisualOdometry() { delete _impl; }
Result VisualOdometry::addFrame(const uint8_t* image, const float* disparity) {
  if (image == nullptr || disparity == nullptr) throw Error("nullptr image/disparity");   // vo.cc:68
  return _impl->addFrame(image, disparity);
}
int VisualOdometry::numPointsAtLevel(int level) const { return _impl->numPointsAtLevel(level); }
const PointVector& VisualOdometry::pointsAtLevel(int level) const { return _impl->pointsAtLevel(level); }
const Trajectory& VisualOdometry::trajectory() const { return _impl->trajectory(); }
bpvo_b200_ctx* VisualOdometry::ctx() const { return _impl->ctx(); }
const bpvo_b200_frame* VisualOdometry::refFrame() const { return _impl->refFrame(); }

}  // namespace bpvo_b200

// -------------------------------------------------------------------------------------------------
// VisualOdometry-level C ABI
// -------------------------------------------------------------------------------------------------
struct bpvo_b200_vo {
  std::unique_ptr<bpvo_b200::VisualOdometry> vo;
  std::unique_ptr<bpvo_b200::PointCloud> last_cloud;
};

int bp_fail(int code, const char* fmt, ...);

#define VO_TRY try {
#define VO_CATCH                                                                                   \
  } catch (const bpvo_b200::Error& e) {                                                           \
    std::string msg = e.what();                                                                    \
    int code = BPVO_B200_ERR_INVALID_ARG;                                                          \
    if (msg.find("no data in frame") != std::string::npos) code = BPVO_B200_ERR_NO_DATA;           \
    else if (msg.find("computeResiduals") != std::string::npos) code = BPVO_B200_ERR_NO_POINTS;    \
    else if (msg.find("CUDA") != std::string::npos || msg.find("cuda") != std::string::npos) code = BPVO_B200_ERR_CUDA; \
    else if (msg.find("not implemented") != std::string::npos || msg.find("not on the accelerated") != std::string::npos) code = BPVO_B200_ERR_UNSUPPORTED; \
    return bp_fail(code, "%s", msg.c_str());                                                       \
  } catch (const std::exception& e) { return bp_fail(BPVO_B200_ERR_INVALID_ARG, "%s", e.what()); }

extern "C" {

int bpvo_b200_vo_create(bpvo_b200_vo** out, const float K[9], float baseline, int rows, int cols, const bpvo_b200_params* p) {
  if (!out || !K || !p) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  VO_TRY
  bpvo_b200::Matrix33 Km; memcpy(Km.m, K, sizeof(Km.m));
  std::unique_ptr<bpvo_b200_vo> h(new bpvo_b200_vo);
  h->vo.reset(new bpvo_b200::VisualOdometry(Km, baseline, bpvo_b200::ImageSize(rows, cols), *p));
  *out = h.release();
  return BPVO_B200_OK;
  VO_CATCH
}
int bpvo_b200_vo_destroy(bpvo_b200_vo* vo) { delete vo; return BPVO_B200_OK; }

int bpvo_b200_vo_add_frame(bpvo_b200_vo* vo, const uint8_t* image, const float* disparity, bpvo_b200_result* out) {
  if (!vo || !out) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  VO_TRY
  bpvo_b200::Result r = vo->vo->addFrame(image, disparity);
  memset(out, 0, sizeof(*out));
  memcpy(out->pose, r.pose.m, sizeof(out->pose));
  out->isKeyFrame = r.isKeyFrame ? 1 : 0; out->keyFramingReason = r.keyFramingReason;
  out->numLevels = (int) r.optimizerStatistics.size();
  for (int i = 0; i < out->numLevels && i < BPVO_B200_MAX_LEVELS; ++i) out->optimizerStatistics[i] = r.optimizerStatistics[i];
  out->numFunEvals = r.numFunEvals;
  out->numPointCloud = r.pointCloud ? (int) r.pointCloud->points.size() : 0;
  vo->last_cloud = std::move(r.pointCloud);
  return BPVO_B200_OK;
  VO_CATCH
}
int bpvo_b200_vo_num_points_at_level(const bpvo_b200_vo* vo, int level, int* n) {
  if (!vo || !n) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  VO_TRY *n = vo->vo->numPointsAtLevel(level); return BPVO_B200_OK; VO_CATCH
}
int bpvo_b200_vo_points_at_level(const bpvo_b200_vo* vo, int level, float* xyzw, int max_points) {
  if (!vo || !xyzw) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  VO_TRY
  const bpvo_b200::PointVector& p = vo->vo->pointsAtLevel(level);
  const size_t n = std::min<size_t>(p.size(), (size_t) std::max(0, max_points));
  if (n) memcpy(xyzw, &p[0].x, n * sizeof(bpvo_b200::Point));
  return BPVO_B200_OK;
  VO_CATCH
}
int bpvo_b200_vo_trajectory(const bpvo_b200_vo* vo, float* poses, int max_poses, int* n) {
  if (!vo || !n) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  const bpvo_b200::Trajectory& t = vo->vo->trajectory();
  *n = (int) t.size();
  for (int i = 0; poses && i < *n && i < max_poses; ++i) memcpy(poses + 16 * (size_t) i, t[i].m, 16 * sizeof(float));
  return BPVO_B200_OK;
}
int bpvo_b200_vo_point_cloud(const bpvo_b200_vo* vo, float* xyzw, float* weights, uint8_t* gray, int max_points, int* n) {
  if (!vo || !n) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  *n = vo->last_cloud ? (int) vo->last_cloud->points.size() : 0;
  for (int i = 0; i < *n && i < max_points; ++i) {
    const bpvo_b200::PointWithInfo& p = vo->last_cloud->points[i];
    if (xyzw) memcpy(xyzw + 4 * (size_t) i, &p.xyzw.x, 16);
    if (weights) weights[i] = p.weight;
    if (gray) gray[i] = p.rgba[0];
  }
  return BPVO_B200_OK;
}
bpvo_b200_ctx* bpvo_b200_vo_ctx(bpvo_b200_vo* vo) { return vo ? vo->vo->ctx() : nullptr; }
const bpvo_b200_frame* bpvo_b200_vo_ref_frame(const bpvo_b200_vo* vo) { return vo ? vo->vo->refFrame() : nullptr; }

}  // extern "C"
