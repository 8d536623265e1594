/*
 * bpvo/vo_pose_estimator.h -- GPU-seam replacement of the reference header of the same name (see bpvo/vo_frame.h here).
 *
 * Same public interface as the reference's VisualOdometryPoseEstimator (bpvo/vo_pose_estimator.h:36-66,
 * bpvo/vo_pose_estimator.cc:55-107); estimatePose() is ONE persistent kernel launch on the B200 (all pyramid levels, all
 * Gauss-Newton iterations, 6x6 solves, convergence tests) behind bpvo_b200_estimate_pose().
 */
#ifndef BPVO_VO_POSE_ESTIMATOR_H
#define BPVO_VO_POSE_ESTIMATOR_H

#include <bpvo/types.h>
// what bpvo/vo.cc picks up through the reference's version of this header (pose_estimator_gn.h -> pose_estimator_base.h -> ...)
#include <bpvo/debug.h>
#include <bpvo/utils.h>
#include <bpvo/math_utils.h>
#include <memory>
#include <vector>

namespace bpvo {

class VisualOdometryFrame;
namespace b200 { struct Group; }

class VisualOdometryPoseEstimator
{
 public:
  VisualOdometryPoseEstimator(const AlgorithmParameters&);
  ~VisualOdometryPoseEstimator();

  /** \return optimizer statistics per pyramid level (entries below maxTestLevel default-constructed) */
  std::vector<OptimizerStatistics>
  estimatePose(const VisualOdometryFrame* ref_frame,
               const VisualOdometryFrame* cur_frame,
               const Matrix44& T_init,
               Matrix44& T_est);

  /** fraction of the last linearize's C*N weights above thresh, counted on the device */
  float getFractionOfGoodPoints(float thresh) const;

  /** weights of the last linearize (channel-major, C*N), downloaded on first use after an estimatePose() */
  const WeightsVector& getWeights() const;

 private:
  AlgorithmParameters _params;
  std::shared_ptr<b200::Group> _group;
  mutable bool _weights_fresh;
  mutable WeightsVector _weights;
}; // VisualOdometryPoseEstimator

}; // bpvo

#endif // BPVO_VO_POSE_ESTIMATOR_H
