// Implementations behind oracle/refstub/opencv2: the three OpenCV calls the hot path makes.
//   cv::pyrDown (CV_8U)            5x5 [1 4 6 4 1]^2, BORDER_REFLECT_101, (sum + 128) >> 8     (bit-exact vs cv2 4.13 golden vectors)
//   cv::GaussianBlur (5x5, CV_32F) separable, symmetric-tap order, no FMA                        (<= 1 ulp vs cv2 4.13 golden vectors)
//   cv::GaussianBlur (3x3, CV_8U)  fixed point 8.8 / 16.16, one rounding                          (bit-exact vs cv2 4.13 golden vectors)
//   cv::Mat::convertTo(CV_32F)     exact
// Test infrastructure only.
#include <opencv2/imgproc/imgproc.hpp>
#include <algorithm>
#include <cmath>
#include <vector>

namespace cv {

static inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * (n - 1) - i; }
  return i;
}

void Mat::convertTo(Mat& dst, int type) const {
  if ((type & 7) == (type_ & 7)) { if (&dst != this) copyTo(dst); return; }
  if ((type_ & 7) != CV_8U || (type & 7) != CV_32F) throw std::logic_error("refstub: convertTo only supports 8U -> 32F");
  Mat out; out.create(rows, cols, CV_32FC1);
  const uint8_t* s = ptr<uint8_t>(); float* d = out.ptr<float>();
  for (size_t i = 0; i < (size_t) rows * cols; ++i) d[i] = (float) s[i];
  dst = out;
}

void pyrDown(const Mat& srcm, Mat& dstm) {
  if ((srcm.type() & 7) != CV_8U) throw std::logic_error("refstub: pyrDown only supports CV_8U");
  const int rows = srcm.rows, cols = srcm.cols, drows = (rows + 1) / 2, dcols = (cols + 1) / 2;
  Mat out; out.create(drows, dcols, CV_8UC1);
  const uint8_t* src = srcm.ptr<uint8_t>(); uint8_t* dst = out.ptr<uint8_t>();
  std::vector<int> rowbuf((size_t) 5 * dcols), xtab((size_t) 5 * dcols);
  for (int x = 0; x < dcols; ++x) for (int k = 0; k < 5; ++k) xtab[(size_t) x * 5 + k] = reflect101(2 * x - 2 + k, cols);
  for (int y = 0; y < drows; ++y) {
    for (int k = 0; k < 5; ++k) {
      const uint8_t* s = src + (size_t) reflect101(2 * y - 2 + k, rows) * cols;
      int* rb = rowbuf.data() + (size_t) k * dcols;
      for (int x = 0; x < dcols; ++x) { const int* xt = &xtab[(size_t) x * 5]; rb[x] = s[xt[0]] + 4 * s[xt[1]] + 6 * s[xt[2]] + 4 * s[xt[3]] + s[xt[4]]; }
    }
    uint8_t* d = dst + (size_t) y * dcols;
    for (int x = 0; x < dcols; ++x) {
      const int v = rowbuf[x] + 4 * rowbuf[dcols + x] + 6 * rowbuf[2 * dcols + x] + 4 * rowbuf[3 * dcols + x] + rowbuf[4 * dcols + x];
      d[x] = (uint8_t) ((v + 128) >> 8);
    }
  }
  dstm = out;
}

// 3x3 on CV_8U (census.cc:65): OpenCV 4.x fixed-point path -- 8-bit taps summing to 256, row pass 8.8, column pass 16.16,
// one rounding (bit-exact vs cv2 4.13 golden vectors)
static void blur3_u8(const Mat& srcm, Mat& dstm, double sigma) {
  const int rows = srcm.rows, cols = srcm.cols;
  const double e = std::exp(-0.5 / (sigma * sigma));
  const int ka = (int) std::lrint(256.0 * (e / (1.0 + 2.0 * e))), kc = 256 - 2 * ka;
  const uint8_t* src = srcm.ptr<uint8_t>();
  std::vector<int> tmp((size_t) rows * cols);
  for (int y = 0; y < rows; ++y) {
    const uint8_t* s = src + (size_t) y * cols;
    for (int x = 0; x < cols; ++x) tmp[(size_t) y * cols + x] = kc * s[x] + ka * (s[reflect101(x - 1, cols)] + s[reflect101(x + 1, cols)]);
  }
  Mat out; out.create(rows, cols, CV_8UC1);
  uint8_t* dst = out.ptr<uint8_t>();
  for (int y = 0; y < rows; ++y) {
    const int* r0 = tmp.data() + (size_t) y * cols;
    const int* rm = tmp.data() + (size_t) reflect101(y - 1, rows) * cols; const int* rp = tmp.data() + (size_t) reflect101(y + 1, rows) * cols;
    for (int x = 0; x < cols; ++x) dst[(size_t) y * cols + x] = (uint8_t) std::min(255, (kc * r0[x] + ka * (rm[x] + rp[x]) + 32768) >> 16);
  }
  dstm = out;
}

// any odd kernel size on CV_32F (cv::Size(): ksize = cvRound(sigma * 8 + 1) | 1, as OpenCV derives it for float images):
// same construction as the 5x5 case below -- getGaussianKernel's formula, separable, reflect-101, symmetric-tap order, no FMA
static void blur_f32_general(const Mat& srcm, Mat& dstm, int ksize, double sigma) {
  const int rows = srcm.rows, cols = srcm.cols, half = ksize / 2;
  std::vector<float> k(ksize);
  { const double sx = sigma > 0 ? sigma : ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8, sc = -0.5 / (sx * sx); double sum = 0;
    for (int i = 0; i < ksize; ++i) { const double x = i - half; k[i] = (float) std::exp(sc * x * x); sum += k[i]; }
    sum = 1.0 / sum; for (int i = 0; i < ksize; ++i) k[i] = (float) (k[i] * sum); }
  const float* src = srcm.ptr<float>();
  std::vector<float> tmp((size_t) rows * cols);
  for (int y = 0; y < rows; ++y) {
    const float* s = src + (size_t) y * cols; float* t = tmp.data() + (size_t) y * cols;
    for (int x = 0; x < cols; ++x) {
      float v = s[x] * k[half];
      for (int j = 1; j <= half; ++j) v = v + (s[reflect101(x - j, cols)] + s[reflect101(x + j, cols)]) * k[half + j];
      t[x] = v;
    }
  }
  Mat out; out.create(rows, cols, CV_32FC1);
  float* dst = out.ptr<float>();
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      float v = k[half] * tmp[(size_t) y * cols + x];
      for (int j = 1; j <= half; ++j) v = v + k[half + j] * (tmp[(size_t) reflect101(y + j, rows) * cols + x] + tmp[(size_t) reflect101(y - j, rows) * cols + x]);
      dst[(size_t) y * cols + x] = v;
    }
  dstm = out;
}

void GaussianBlur(const Mat& srcm, Mat& dstm, Size ksize, double sigmaX, double) {
  if ((srcm.type() & 7) == CV_8U && ksize.width == 3 && ksize.height == 3 && sigmaX > 0) { blur3_u8(srcm, dstm, sigmaX); return; }
  if ((srcm.type() & 7) == CV_32F && ksize.width <= 0 && sigmaX > 0) ksize = Size(((int) std::lrint(sigmaX * 8 + 1)) | 1, ((int) std::lrint(sigmaX * 8 + 1)) | 1);
  if ((srcm.type() & 7) == CV_32F && ksize.width == ksize.height && (ksize.width & 1) && ksize.width != 5) { blur_f32_general(srcm, dstm, ksize.width, sigmaX); return; }
  if ((srcm.type() & 7) != CV_32F || ksize.width != 5 || ksize.height != 5)
    throw std::logic_error("refstub: GaussianBlur supports odd square kernels on CV_32F and 3x3 on CV_8U");
  const int rows = srcm.rows, cols = srcm.cols;
  float k[5];
  { const double s = sigmaX > 0 ? sigmaX : ((5 - 1) * 0.5 - 1) * 0.3 + 0.8; const double sc = -0.5 / (s * s); double sum = 0;
    for (int i = 0; i < 5; ++i) { const double x = i - 2.0; k[i] = (float) std::exp(sc * x * x); sum += k[i]; }
    sum = 1.0 / sum; for (int i = 0; i < 5; ++i) k[i] = (float) (k[i] * sum); }
  const float k0 = k[2], k1 = k[3], k2 = k[4];
  const float* src = srcm.ptr<float>();
  std::vector<float> tmp((size_t) rows * cols);
  // row pass: borders with reflect-101 indexing, interior as a plain (auto-vectorisable) loop; same expression everywhere
  for (int y = 0; y < rows; ++y) {
    const float* s = src + (size_t) y * cols; float* t = tmp.data() + (size_t) y * cols;
    auto at = [&](int x) {
      const float a = s[reflect101(x - 1, cols)] + s[reflect101(x + 1, cols)], b = s[reflect101(x - 2, cols)] + s[reflect101(x + 2, cols)];
      float v = s[x] * k0; v = v + a * k1; v = v + b * k2; return v;
    };
    const int lo = std::min(2, cols), hi = std::max(lo, cols - 2);
    for (int x = 0; x < lo; ++x) t[x] = at(x);
    for (int x = lo; x < hi; ++x) { float v = s[x] * k0; v = v + (s[x - 1] + s[x + 1]) * k1; v = v + (s[x - 2] + s[x + 2]) * k2; t[x] = v; }
    for (int x = hi; x < cols; ++x) t[x] = at(x);
  }
  Mat out; out.create(rows, cols, CV_32FC1);
  float* dst = out.ptr<float>();
  for (int y = 0; y < rows; ++y) {
    const float* r0 = tmp.data() + (size_t) y * cols;
    const float* rm1 = tmp.data() + (size_t) reflect101(y - 1, rows) * cols; const float* rp1 = tmp.data() + (size_t) reflect101(y + 1, rows) * cols;
    const float* rm2 = tmp.data() + (size_t) reflect101(y - 2, rows) * cols; const float* rp2 = tmp.data() + (size_t) reflect101(y + 2, rows) * cols;
    float* d = dst + (size_t) y * cols;
    for (int x = 0; x < cols; ++x) { float v = k0 * r0[x]; v = v + k1 * (rp1[x] + rm1[x]); v = v + k2 * (rp2[x] + rm2[x]); d[x] = v; }
  }
  dstm = out;
}

}  // namespace cv
