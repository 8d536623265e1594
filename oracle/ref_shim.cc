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
