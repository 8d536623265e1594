// seam_test_shim.cc -- C entry points over bpvo::VisualOdometry for the drop-in test: the reference's vo.cc compiled
// against the GPU seam (tests/test_gpu_reference_seam.py loads this through ctypes).  Test infrastructure: same function
// names and POD layouts as the VisualOdometry part of oracle/ref_shim.cc, so that one Python wrapper drives both libraries.
#include <bpvo/point_cloud.h>
#include <bpvo/trajectory.h>
#include <bpvo/vo.h>

#include <cstring>
#include <memory>
#include <string>

#include "bpvo_oracle.h"     // POD layouts (orc_params, orc_result)

using namespace bpvo;

namespace {
thread_local std::string g_err;
struct Vo { std::unique_ptr<VisualOdometry> vo; std::unique_ptr<PointCloud> cloud; };
}

#define TRY try {
#define CATCH(rv) } catch (const std::exception& e) { g_err = e.what(); return rv; }

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

void* ref_vo_create(const float K[9], float baseline, int rows, int cols, const orc_params* q) {
  TRY
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
  Matrix33 Km; memcpy(Km.data(), K, 9 * sizeof(float));
  std::unique_ptr<Vo> h(new Vo);
  h->vo.reset(new VisualOdometry(Km, baseline, ImageSize(rows, cols), p));
  return h.release();
  CATCH(nullptr)
}
void ref_vo_destroy(void* h) { delete (Vo*) h; }
int ref_vo_add_frame(void* hh, const uint8_t* image, const float* disparity, orc_result* out) {
  TRY
  Vo* h = (Vo*) hh;
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
  CATCH(-1)
}
int ref_vo_num_points_at_level(void* hh, int level) { TRY return ((Vo*) hh)->vo->numPointsAtLevel(level); CATCH(-1) }
int ref_vo_trajectory(void* hh, float* poses, int max_poses) {
  const Trajectory& t = ((Vo*) hh)->vo->trajectory();
  const int n = (int) t.size();
  for (int i = 0; i < n && i < max_poses; ++i) memcpy(poses + 16 * (size_t) i, t[i].data(), 16 * sizeof(float));
  return n;
}
int ref_vo_point_cloud(void* hh, float* xyzw, float* weights, uint8_t* gray, int max_points) {
  Vo* h = (Vo*) hh;
  const int n = h->cloud ? (int) h->cloud->size() : 0;
  for (int i = 0; i < n && i < max_points; ++i) {
    const PointWithInfo& p = (*h->cloud)[i];
    memcpy(xyzw + 4 * (size_t) i, p.xyzw().data(), 16); weights[i] = p.weight(); gray[i] = p.rgba()[0];
  }
  return n;
}

}  // extern "C"
