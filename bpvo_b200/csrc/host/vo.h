// host/vo.h -- C++ host shim with the reference's public interface (bpvo/vo.h:33-107) on top of the
// seam-level C ABI.  It is bpvo/vo.cc's VisualOdometry::Impl restated against bpvo_b200_frame /
// bpvo_b200_ctx instead of VisualOdometryFrame / VisualOdometryPoseEstimator: the key-framing state
// machine stays on the host, everything below it runs on the GPU.
//
// Matrix types: the reference uses Eigen (Matrix33, Matrix44, column-major).  Eigen is not part of
// this repository's toolchain, so the shim exposes POD column-major matrices with the same memory
// layout; with Eigen available, `Eigen::Map<Matrix44>(m.data())` / `Matrix44::data()` convert for free.
#pragma once

#include <stdint.h>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/bpvo_b200.h"

namespace bpvo_b200 {

struct Error : public std::logic_error {       // bpvo::Error (bpvo/utils.h:211-220)
  explicit Error(const std::string& what) : std::logic_error(what) {}
};

struct Matrix33 { float m[9]; float* data() { return m; } const float* data() const { return m; }
                  float& operator()(int r, int c) { return m[c * 3 + r]; } float operator()(int r, int c) const { return m[c * 3 + r]; } };
struct Matrix44 { float m[16]; float* data() { return m; } const float* data() const { return m; }
                  float& operator()(int r, int c) { return m[c * 4 + r]; } float operator()(int r, int c) const { return m[c * 4 + r]; }
                  static Matrix44 Identity(); };
struct Point { float x, y, z, w; };
typedef std::vector<Point> PointVector;

struct ImageSize { int rows = 0, cols = 0; ImageSize(int r = 0, int c = 0) : rows(r), cols(c) {} };   // bpvo/types.h:568-577

typedef bpvo_b200_params AlgorithmParameters;    // POD mirror of bpvo::AlgorithmParameters
typedef bpvo_b200_stats OptimizerStatistics;

struct PointWithInfo { Point xyzw; uint8_t rgba[4]; float weight; uint8_t pad[8]; };   // bpvo/point_cloud.h:50-58 (32 bytes)
struct PointCloud { std::vector<PointWithInfo> points; Matrix44 pose; };

struct Result {                                  // bpvo::Result (bpvo/types.h:496-566)
  Matrix44 pose;
  float covariance[36];                          // identity: never computed by the reference (types.cc:324)
  std::vector<OptimizerStatistics> optimizerStatistics;
  bool isKeyFrame = false;
  int keyFramingReason = BPVO_B200_KF_NONE;
  std::unique_ptr<PointCloud> pointCloud;
  int numFunEvals = 0;
  Result();
};

class Trajectory {                               // bpvo/trajectory.{h,cc}
 public:
  void push_back(const Matrix44& T);
  const Matrix44& back() const { return _poses.back(); }
  size_t size() const { return _poses.size(); }
  const Matrix44& operator[](size_t i) const { return _poses[i]; }
 private:
  std::vector<Matrix44> _poses;
};

class VisualOdometry {                           // bpvo::VisualOdometry (bpvo/vo.h:33-107)
 public:
  VisualOdometry(const Matrix33& K, float baseline, ImageSize, const AlgorithmParameters&);
  ~VisualOdometry();
  VisualOdometry(const VisualOdometry&) = delete;
  VisualOdometry& operator=(const VisualOdometry&) = delete;
  Result addFrame(const uint8_t* image, const float* disparity);
  int numPointsAtLevel(int level = -1) const;
  const PointVector& pointsAtLevel(int level = -1) const;
  const Trajectory& trajectory() const;
  bpvo_b200_ctx* ctx() const;
  const bpvo_b200_frame* refFrame() const;
 private:
  class Impl;
  Impl* _impl;
};

}  // namespace bpvo_b200
