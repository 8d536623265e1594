/*
 * bpvo_oracle.cc -- dependency-free CPU restatement of halismai/bpvo's per-frame Gauss-Newton
 * dense-alignment path (reference @ 343d9da).  TEST INFRASTRUCTURE ONLY (see bpvo_oracle.h).
 *
 * PARITY STATUS: pinned bit-exact against the reference's own hot-path sources (23 .cc files compiled from
 * /root/reference against the stand-in Eigen/OpenCV headers of oracle/refstub -> oracle/_ref/libbpvo_ref.so):
 * every stage, PoseEstimatorGN::linearize, and whole VisualOdometry::addFrame streams incl. per-level iteration
 * counts and key-frame decisions (tests/test_oracle_vs_reference.py).  cv::pyrDown / cv::GaussianBlur are
 * pinned against cv2 4.13 golden vectors (tests/golden/).  The reference's own tests pin nothing here.
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 * The reference's SSE/AVX intrinsics are kept where it has them; its documented quirks
 *  are reproduced on purpose.
 *
 * Build: see oracle/Makefile (-O3 -msse4.1 -mavx -mfpmath=sse, no FMA contraction -- the flags the
 * reference builds itself with, cmake/SetCompilerOptions.cmake:53-58).
 */
#include "bpvo_oracle.h"

#include <immintrin.h>
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#if defined(_OPENMP)
#include <omp.h>
#endif

namespace {

thread_local std::string g_last_error;

// ------------------------------------------------------------------------------------------
// aligned storage (bpvo/aligned_allocator.h, types.h:56-73: DefaultAlignment 32 with AVX)
// ------------------------------------------------------------------------------------------
template <typename T> struct AlignedAlloc {
  typedef T value_type;
  AlignedAlloc() {}
  template <class U> AlignedAlloc(const AlignedAlloc<U>&) {}
  T* allocate(size_t n) {
    void* p = nullptr;
    if (posix_memalign(&p, 64, std::max<size_t>(n * sizeof(T), 64)) != 0) throw std::bad_alloc();
    return static_cast<T*>(p);
  }
  void deallocate(T* p, size_t) { free(p); }
  template <class U> bool operator==(const AlignedAlloc<U>&) const { return true; }
  template <class U> bool operator!=(const AlignedAlloc<U>&) const { return false; }
};
template <typename T> using avec = std::vector<T, AlignedAlloc<T>>;

// ------------------------------------------------------------------------------------------
// small fixed-size matrices, column-major like Eigen
// ------------------------------------------------------------------------------------------
struct M44 {
  float m[16];
  float& operator()(int r, int c) { return m[c * 4 + r]; }
  float operator()(int r, int c) const { return m[c * 4 + r]; }
  static M44 Identity() {
    M44 r; std::fill_n(r.m, 16, 0.0f);
    r(0,0) = r(1,1) = r(2,2) = r(3,3) = 1.0f; return r;
  }
};
struct M33 {
  float m[9];
  float& operator()(int r, int c) { return m[c * 3 + r]; }
  float operator()(int r, int c) const { return m[c * 3 + r]; }
};
struct M34 {
  float m[12];
  float& operator()(int r, int c) { return m[c * 3 + r]; }
  float operator()(int r, int c) const { return m[c * 3 + r]; }
};

static M44 mul(const M44& a, const M44& b) {
  M44 r;
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 4; ++i) {
      float s = a(i,0) * b(0,j);
      for (int k = 1; k < 4; ++k) s += a(i,k) * b(k,j);
      r(i,j) = s;
    }
  return r;
}

// general 4x4 inverse in float (Eigen Matrix4f::inverse(), cofactor expansion)
static M44 inverse(const M44& a) {
  const float* m = a.m;
  float inv[16];
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
  float det = m[0]*inv[0] + m[1]*inv[4] + m[2]*inv[8] + m[3]*inv[12];
  float idet = 1.0f / det;
  M44 r;
  for (int i = 0; i < 16; ++i) r.m[i] = inv[i] * idet;
  return r;
}

// math::TwistToMatrix<float> (bpvo/math_utils.h:140-168); p[0:3]=omega, p[3:6]=v
static M44 twist_to_matrix(const float p[6]) {
  M44 ret = M44::Identity();
  const float theta = std::sqrt(p[0]*p[0] + p[1]*p[1] + p[2]*p[2]);
  if (theta > 1e-8) {
    // `T a = ::sin(theta)` with T = float: with libstdc++ >= 6 and <math.h> included (OpenCV's types_c.h does) the global
    // ::sin / ::cos resolve to the float overloads; `1.0 - ::cos(theta)` and `1.0 / theta` are evaluated in double and
    // narrowed.  (A pre-GCC-6 toolchain would call the double versions: <= 1 ulp difference in a, b.)
    float a = sinf(theta);
    float b = (float) (1.0 - (double) cosf(theta));
    float t_i = (float) (1.0 / (double) theta);
    // S = t_i * skew(w)
    float S[3][3] = {{0, -p[2]*t_i, p[1]*t_i}, {p[2]*t_i, 0, -p[0]*t_i}, {-p[1]*t_i, p[0]*t_i, 0}};
    S[0][1] = t_i * (-p[2]); S[0][2] = t_i * p[1]; S[1][0] = t_i * p[2];
    S[1][2] = t_i * (-p[0]); S[2][0] = t_i * (-p[1]); S[2][1] = t_i * p[0];
    float S2[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
      float s = S[i][0]*S[0][j]; s += S[i][1]*S[1][j]; s += S[i][2]*S[2][j]; S2[i][j] = s;
    }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
      ret(i,j) = ((i == j) ? 1.0f : 0.0f) + a*S[i][j] + b*S2[i][j];
    const float c1 = b * t_i, c2 = (theta - a) * t_i;
    for (int i = 0; i < 3; ++i) {
      float s = 0.0f;
      for (int j = 0; j < 3; ++j) {
        float v = ((i == j) ? 1.0f : 0.0f) + c1*S[i][j] + c2*S2[i][j];
        s = (j == 0) ? v * p[3] : s + v * p[3+j];
      }
      ret(i,3) = s;
    }
  } else {
    ret(0,3) = p[3]; ret(1,3) = p[4]; ret(2,3) = p[5];
  }
  return ret;
}

// ------------------------------------------------------------------------------------------
// Eigen::LDLT (3.2.x, README.md:20 pins 3.2.8) restated: pivoted in-place LDL^T, lower storage
// ------------------------------------------------------------------------------------------
template <typename S, int N> struct LDLT {
  S a[N][N];   // lower triangle holds L (unit diag) and D
  int tr[N];
  void compute(const S* Hcolmajor) {
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) a[i][j] = Hcolmajor[j*N + i];
    S cutoff = 0, temp[N];
    for (int k = 0; k < N; ++k) {
      int idx = k; S big = std::abs(a[k][k]);
      for (int i = k + 1; i < N; ++i) if (std::abs(a[i][i]) > big) { big = std::abs(a[i][i]); idx = i; }
      if (k == 0) cutoff = std::abs(std::numeric_limits<S>::epsilon() * big);
      if (big < cutoff) { for (int i = k; i < N; ++i) tr[i] = i; break; }
      tr[k] = idx;
      if (k != idx) {
        int s = N - idx - 1;
        for (int j = 0; j < k; ++j) std::swap(a[k][j], a[idx][j]);
        for (int i = 0; i < s; ++i) std::swap(a[N - s + i][k], a[N - s + i][idx]);
        std::swap(a[k][k], a[idx][idx]);
        for (int i = k + 1; i < idx; ++i) { S t = a[i][k]; a[i][k] = a[idx][i]; a[idx][i] = t; }
      }
      int rs = N - k - 1;
      if (k > 0) {
        for (int j = 0; j < k; ++j) temp[j] = a[j][j] * a[k][j];
        S s = 0; for (int j = 0; j < k; ++j) s += a[k][j] * temp[j];
        a[k][k] -= s;
        for (int i = 0; i < rs; ++i) {
          S t = 0; for (int j = 0; j < k; ++j) t += a[k+1+i][j] * temp[j];
          a[k+1+i][k] -= t;
        }
      }
      if (rs > 0 && std::abs(a[k][k]) > cutoff)
        for (int i = 0; i < rs; ++i) a[k+1+i][k] /= a[k][k];
    }
  }
  void solve(const S* b, S* x) const {
    for (int i = 0; i < N; ++i) x[i] = b[i];
    for (int k = 0; k < N; ++k) if (tr[k] != k) std::swap(x[k], x[tr[k]]);
    for (int i = 0; i < N; ++i) for (int j = 0; j < i; ++j) x[i] -= a[i][j] * x[j];
    S dmax = 0; for (int i = 0; i < N; ++i) dmax = std::max(dmax, std::abs(a[i][i]));
    S tol = std::max(dmax * std::numeric_limits<S>::epsilon(), S(1) / std::numeric_limits<S>::max());
    for (int i = 0; i < N; ++i) { if (std::abs(a[i][i]) > tol) x[i] /= a[i][i]; else x[i] = 0; }
    for (int i = N - 1; i >= 0; --i) for (int j = i + 1; j < N; ++j) x[i] -= a[j][i] * x[j];
    for (int k = N - 1; k >= 0; --k) if (tr[k] != k) std::swap(x[k], x[tr[k]]);
  }
};

// Eigen isApprox for vectors: ||a-b||^2 <= prec^2 * min(||a||^2,||b||^2)
template <typename S> static bool is_approx6(const S* a, const S* b, S prec) {
  S d = 0, na = 0, nb = 0;
  for (int i = 0; i < 6; ++i) { d += (a[i]-b[i])*(a[i]-b[i]); na += a[i]*a[i]; nb += b[i]*b[i]; }
  return d <= prec * prec * std::min(na, nb);
}

// PoseEstimatorData_::solve / solve2Augmented (bpvo/pose_estimator_base.h:90-111, 136-148)
static bool solve6(const float* H, const float* G, float* dp) {
  LDLT<float, 6> l; l.compute(H); l.solve(G, dp);
  float Hd[6];
  for (int i = 0; i < 6; ++i) { float s = H[0*6+i]*dp[0]; for (int k = 1; k < 6; ++k) s += H[k*6+i]*dp[k]; Hd[i] = s; }
  bool ok = is_approx6<float>(Hd, G, 1e-5f);
  if (!ok) {
    float dmax = H[0]; for (int i = 1; i < 6; ++i) dmax = std::max(dmax, H[i*6+i]);
    double u = 0.001 * dmax;
    double Hq[36], Gq[6], dq[6];
    for (int i = 0; i < 36; ++i) Hq[i] = H[i];
    for (int i = 0; i < 6; ++i) { Gq[i] = G[i]; Hq[i*6+i] += u; }
    LDLT<double, 6> l2; l2.compute(Hq); l2.solve(Gq, dq);
    double Hd2[6];
    for (int i = 0; i < 6; ++i) { double s = Hq[0*6+i]*dq[0]; for (int k = 1; k < 6; ++k) s += Hq[k*6+i]*dq[k]; Hd2[i] = s; }
    ok = is_approx6<double>(Hd2, Gq, 1e-12);
    for (int i = 0; i < 6; ++i) dp[i] = (float) dq[i];
  }
  return ok;
}

// ------------------------------------------------------------------------------------------
// image ops
// ------------------------------------------------------------------------------------------
static inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * (n - 1) - i; }
  return i;
}

// cv::pyrDown, CV_8U, 5x5 [1 4 6 4 1], BORDER_REFLECT_101, (sum+128)>>8 (bpvo/image_pyramid.cc:49;
// OpenCV imgproc/pyramids.cpp pyrDown_).  Third-party arithmetic pinned by tests/golden/pyrdown_*.npz.
static void pyr_down(const uint8_t* src, int rows, int cols, uint8_t* dst) {
  const int drows = (rows + 1) / 2, dcols = (cols + 1) / 2;
  std::vector<int> rowbuf((size_t) 5 * dcols);
  std::vector<int> xtab((size_t) 5 * dcols);
  for (int x = 0; x < dcols; ++x)
    for (int k = 0; k < 5; ++k) xtab[(size_t) x * 5 + k] = reflect101(2 * x - 2 + k, cols);
  for (int y = 0; y < drows; ++y) {
    for (int k = 0; k < 5; ++k) {
      const uint8_t* s = src + (size_t) reflect101(2 * y - 2 + k, rows) * cols;
      int* rb = rowbuf.data() + (size_t) k * dcols;
      for (int x = 0; x < dcols; ++x) {
        const int* xt = &xtab[(size_t) x * 5];
        rb[x] = s[xt[0]] + 4 * s[xt[1]] + 6 * s[xt[2]] + 4 * s[xt[3]] + s[xt[4]];
      }
    }
    uint8_t* d = dst + (size_t) y * dcols;
    for (int x = 0; x < dcols; ++x) {
      int v = rowbuf[x] + 4 * rowbuf[dcols + x] + 6 * rowbuf[2*dcols + x] + 4 * rowbuf[3*dcols + x] + rowbuf[4*dcols + x];
      d[x] = (uint8_t) ((v + 128) >> 8);
    }
  }
}

// cv::getGaussianKernel(5, sigma, CV_32F) (OpenCV imgproc/smooth.cpp): exp in double, stored float,
// normalised by the double sum of the float taps.
static void gaussian_kernel5(float sigma, float k[5]) {
  double sigmaX = sigma > 0 ? (double) sigma : ((5 - 1) * 0.5 - 1) * 0.3 + 0.8;
  double scale2X = -0.5 / (sigmaX * sigmaX);
  double sum = 0;
  for (int i = 0; i < 5; ++i) { double x = i - 2.0; k[i] = (float) std::exp(scale2X * x * x); sum += k[i]; }
  sum = 1.0 / sum;
  for (int i = 0; i < 5; ++i) k[i] = (float) (k[i] * sum);
}

// cv::GaussianBlur(5x5) on CV_32F, BORDER_REFLECT_101, separable, symmetric-tap evaluation order of
// OpenCV's SymmRowSmallFilter / SymmColumnFilter: (s0*k0 + (s-1+s1)*k1) + (s-2+s2)*k2, float, no FMA
// (bpvo/bitplanes_descriptor.cc:55-56).  Pinned by tests/golden/gaussblur5_*.npz to 1e-6.
static void gaussian_blur5(const float* src, int rows, int cols, float sigma, float* dst) {
  float k[5]; gaussian_kernel5(sigma, k);
  const float k0 = k[2], k1 = k[3], k2 = k[4];
  std::vector<float> tmp((size_t) rows * cols);
  for (int y = 0; y < rows; ++y) {
    const float* s = src + (size_t) y * cols; float* t = tmp.data() + (size_t) y * cols;
    for (int x = 0; x < cols; ++x) {
      float a = s[reflect101(x-1, cols)] + s[reflect101(x+1, cols)];
      float b = s[reflect101(x-2, cols)] + s[reflect101(x+2, cols)];
      float v = s[x] * k0; v = v + a * k1; v = v + b * k2; t[x] = v;
    }
  }
  for (int y = 0; y < rows; ++y) {
    const float* r0 = tmp.data() + (size_t) y * cols;
    const float* rm1 = tmp.data() + (size_t) reflect101(y-1, rows) * cols;
    const float* rp1 = tmp.data() + (size_t) reflect101(y+1, rows) * cols;
    const float* rm2 = tmp.data() + (size_t) reflect101(y-2, rows) * cols;
    const float* rp2 = tmp.data() + (size_t) reflect101(y+2, rows) * cols;
    float* d = dst + (size_t) y * cols;
    for (int x = 0; x < cols; ++x) {
      float v = k0 * r0[x]; v = v + k1 * (rp1[x] + rm1[x]); v = v + k2 * (rp2[x] + rm2[x]); d[x] = v;
    }
  }
}

// cv::GaussianBlur on CV_32F with an arbitrary odd kernel size (gradient_descriptor.cc:52-53 with cv::Size() -> OpenCV derives
// ksize = cvRound(sigma * 8 + 1) | 1 for float images; imgproc.cc:166-171 imsmooth: ksize = max(5, 2 round(sigma) + 1)).
// Same construction as gaussian_blur5: taps from cv::getGaussianKernel's formula, separable, BORDER_REFLECT_101, symmetric-tap
// order s0 k0 + (s-1 + s1) k1 + (s-2 + s2) k2 + ..., float, no FMA.  THIRD-PARTY arithmetic: within a few ulp of cv2 4.13
// (tests/test_oracle_cpu.py), bit-identical to the stand-in cv::GaussianBlur of oracle/refstub that oracle/_ref links.
static void gaussian_blur_f32(const float* src, int rows, int cols, int ksize, double sigma, float* dst) {
  if (ksize <= 0) ksize = ((int) std::lrint(sigma * 8 + 1)) | 1;
  if (ksize == 5) { gaussian_blur5(src, rows, cols, (float) sigma, dst); return; }
  const int half = ksize / 2;
  std::vector<float> k(ksize);
  { const double sx = sigma > 0 ? sigma : ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8, sc = -0.5 / (sx * sx); double sum = 0;
    for (int i = 0; i < ksize; ++i) { const double x = i - half; k[i] = (float) std::exp(sc * x * x); sum += k[i]; }
    sum = 1.0 / sum; for (int i = 0; i < ksize; ++i) k[i] = (float) (k[i] * sum); }
  std::vector<float> tmp((size_t) rows * cols);
  for (int y = 0; y < rows; ++y) {
    const float* s = src + (size_t) y * cols; float* t = tmp.data() + (size_t) y * cols;
    for (int x = 0; x < cols; ++x) {
      float v = s[x] * k[half];
      for (int j = 1; j <= half; ++j) v = v + (s[reflect101(x - j, cols)] + s[reflect101(x + j, cols)]) * k[half + j];
      t[x] = v;
    }
  }
  for (int y = 0; y < rows; ++y) {
    float* d = dst + (size_t) y * cols;
    for (int x = 0; x < cols; ++x) {
      float v = k[half] * tmp[(size_t) y * cols + x];
      for (int j = 1; j <= half; ++j) v = v + k[half + j] * (tmp[(size_t) reflect101(y + j, rows) * cols + x] + tmp[(size_t) reflect101(y - j, rows) * cols + x]);
      d[x] = v;
    }
  }
}
// imsmooth (imgproc.cc:166-171)
static void imsmooth(const float* src, int rows, int cols, double sigma, float* dst) {
  const int k = std::max(5, 2 * (int) std::round(sigma) + 1);
  gaussian_blur_f32(src, rows, cols, k, sigma, dst);
}
// xgradient / ygradient (imgproc.h:215-266): 0.5 * central difference, one-sided 0.5 * (I1 - I0) on the first / last column / row
static void xgradient(const float* I, int rows, int cols, float* Ix) {
  for (int y = 0; y < rows; ++y) {
    const float* s = I + (size_t) y * cols; float* d = Ix + (size_t) y * cols;
    d[0] = 0.5f * (s[1] - s[0]);
    for (int x = 1; x < cols - 1; ++x) d[x] = 0.5f * (s[x + 1] - s[x - 1]);
    d[cols - 1] = 0.5f * (s[cols - 1] - s[cols - 2]);
  }
}
static void ygradient(const float* I, int rows, int cols, float* Iy) {
  for (int x = 0; x < cols; ++x) Iy[x] = 0.5f * (I[cols + x] - I[x]);
  for (int y = 1; y < rows - 1; ++y)
    for (int x = 0; x < cols; ++x) Iy[(size_t) y * cols + x] = 0.5f * (I[(size_t) (y + 1) * cols + x] - I[(size_t) (y - 1) * cols + x]);
  for (int x = 0; x < cols; ++x) Iy[(size_t) (rows - 1) * cols + x] = 0.5f * (I[(size_t) (rows - 1) * cols + x] - I[(size_t) (rows - 2) * cols + x]);
}

// census (bpvo/census.cc:42-91, v128.h:87-120): bit k set iff neighbour_k >= centre (unsigned),
// k: 0=NW 1=N 2=NE 3=W 4=E 5=SW 6=S 7=SE; first/last row and column are 0.  sigma>0 pre-blur
// (census.cc:64-65, a third-party u8 3x3 GaussianBlur) is gaussian_blur3_u8() below, applied by compute_descriptor().
static void census(const uint8_t* src, int rows, int cols, uint8_t* dst) {
  memset(dst, 0, (size_t) rows * cols);
  for (int y = 1; y < rows - 1; ++y) {
    const uint8_t* s = src + (size_t) y * cols; uint8_t* d = dst + (size_t) y * cols;
    for (int x = 1; x < cols - 1; ++x) {
      const uint8_t c = s[x];
      d[x] = (uint8_t) (((s[x-cols-1] >= c) << 0) | ((s[x-cols] >= c) << 1) | ((s[x-cols+1] >= c) << 2) |
                        ((s[x-1] >= c) << 3) | ((s[x+1] >= c) << 4) |
                        ((s[x+cols-1] >= c) << 5) | ((s[x+cols] >= c) << 6) | ((s[x+cols+1] >= c) << 7));
    }
  }
}

// gradientAbsMag, 4 lanes (bpvo/imgproc.cc:33-43), emulated lane by lane
static inline float grad_abs_mag(const float* src, int stride) {
  float ix = std::fabs(src[-1] - src[1]);
  float iy = std::fabs(src[-stride] - src[stride]);
  return ix + iy;
}

// gradientAbsoluteMagnitude (bpvo/imgproc.cc:45-78) -- literal, incl. the scalar-tail "+" (Q4),
// the dst[cols] write and the col-0 wrap-around read.
static void saliency_first(const float* src_ptr, int rows, int cols, float* dst_ptr) {
  std::fill_n(dst_ptr, cols, 0.0f);
  const float* src = src_ptr + cols; float* dst = dst_ptr + cols;
  const int n = cols & ~3;
  for (int r = 2; r < rows; ++r) {
    int x = 0;
    for (; x < n; x += 4) for (int l = 0; l < 4; ++l) dst[x + l] = grad_abs_mag(src + x + l, cols);
    for (; x < cols; ++x) dst[x] = std::fabs(src[x+1] - src[x-1]) + std::fabs(src[x+cols] + src[x-cols]);
    dst[x] = 0.0f;          // x == cols: next row's column 0
    dst[cols-1] = 0.0f;
    dst += cols; src += cols;
  }
  std::fill_n(dst, cols, 0.0f);
}

// gradientAbsoluteMagnitudeAcc (bpvo/imgproc.cc:104-127) -- literal, incl. the store-to-dst bug (Q3)
static void saliency_acc(const float* src, int rows, int cols, float* dst) {
  const int n = cols & ~3;
  src += cols; dst += cols;
  for (int r = 2; r < rows; ++r, src += cols, dst += cols) {
    int x = 0;
    for (; x < n; x += 4) {
      float g[4];
      for (int l = 0; l < 4; ++l) g[l] = dst[x + l] + grad_abs_mag(src + x + l, cols);
      for (int l = 0; l < 4; ++l) dst[l] = g[l];     // sic: stored at dst, not dst+x
    }
    for (; x < cols; ++x) dst[x] += std::fabs(src[x-1] - src[x+1]) + std::fabs(src[x-cols] + src[x+cols]);
    dst[x] = 0.0f;
    dst[cols-1] = 0.0f;
  }
}

// DenseDescriptor::computeSaliencyMap (bpvo/dense_descriptor.cc:92-100) and the intensity override
// (intensity_descriptor.cc:45-53, same first function)
static void saliency_map(const float* planes, int channels, int rows, int cols, float* dst) {
  // reads of src[-1] at (row 1, col 0) and src[+1]/dst[cols] at the last processed row stay inside
  // the rows x cols buffers, exactly as in the reference.
  saliency_first(planes, rows, cols, dst);
  for (int c = 1; c < channels; ++c) saliency_acc(planes + (size_t) c * rows * cols, rows, cols, dst);
}

// IsLocalMax<float> (bpvo/imgproc.h:93-165), WITH_SIMD variant for radius 1 (3x4 window, Q5)
struct IsLocalMax {
  const float* ptr; int stride, radius;
  bool operator()(int row, int col) const {
    if (radius <= 0) return true;
    if (radius == 1) {
      const float* p = ptr + (size_t) row * stride + col;
      const float v = *p;
      // masks 13/15/15: row 0 lanes (p-1, p, p+1, p+2) -> 1,0,1,1 ; rows +-1 all four
      bool r0 = (v > p[-1]) && !(v > p[0]) && (v > p[1]) && (v > p[2]);
      const float* pu = p - 1 - stride; const float* pd = p - 1 + stride;
      bool ru = (v > pu[0]) && (v > pu[1]) && (v > pu[2]) && (v > pu[3]);
      bool rd = (v > pd[0]) && (v > pd[1]) && (v > pd[2]) && (v > pd[3]);
      return r0 && ru && rd;
    }
    const float v = ptr[(size_t) row * stride + col];
    for (int r = -radius; r <= radius; ++r)
      for (int c = -radius; c <= radius; ++c)
        if (!(!r && !c) && ptr[(size_t) (row + r) * stride + col + c] >= v) return false;
    return true;
  }
};

// ------------------------------------------------------------------------------------------
// RigidBodyWarp (bpvo/rigid_body_warp.h:29-155, rigid_body_warp.cc:47-315)
// ------------------------------------------------------------------------------------------
struct Warp {
  M33 K; float b; M34 P; M44 T, T_inv; bool use_rcp = true;

  void init(const M33& K_, float b_) { K = K_; b = b_; T = M44::Identity(); T_inv = M44::Identity(); }

  // makePoint (rigid_body_warp.h:47-60); Z = Bf * (1.0/d) in double (Q12)
  void make_point(float x, float y, float d, float* out) const {
    float fx = K(0,0), fy = K(1,1), cx = K(0,2), cy = K(1,2);
    float Bf = b * fx;
    float Z = (float) ((double) Bf * (1.0 / (double) d));
    float X = (x - cx) * Z * (1.0f / fx);
    float Y = (y - cy) * Z * (1.0f / fy);
    out[0] = X; out[1] = Y; out[2] = Z; out[3] = 1.0f;
  }

  void set_normalization(const M44& Tn) { T = Tn; T_inv = inverse(Tn); }

  // setPose: P = K * T[0:3,:] (rigid_body_warp.h:111-114)
  void set_pose(const M44& pose) {
    for (int j = 0; j < 4; ++j) for (int i = 0; i < 3; ++i) {
      float s = K(i,0) * pose(0,j); s += K(i,1) * pose(1,j); s += K(i,2) * pose(2,j); P(i,j) = s;
    }
  }

  // paramsToPose = T_inv * exp(p) * T (rigid_body_warp.h:130-138)
  M44 params_to_pose(const float p[6]) const { return mul(mul(T_inv, twist_to_matrix(p)), T); }

  // scalar jacobian (rigid_body_warp.h:94-106)
  void jacobian(const float* p, float Ix, float Iy, float* J) const {
    float X = p[0], Y = p[1], Z = p[2];
    float fx = K(0,0), fy = K(1,1);
    float s = T(0,0), c1 = T_inv(0,3), c2 = T_inv(1,3), c3 = T_inv(2,3);
    J[0] = -1.0f/(Z*Z)*(Ix*X*fx+Iy*Y*fy)*(Y-c2)-(Iy*fy*(Z-c3))/Z;
    J[1] = 1.0f/(Z*Z)*(Ix*X*fx+Iy*Y*fy)*(X-c1)+(Ix*fx*(Z-c3))/Z;
    J[2] = (Iy*fy*(X-c1))/Z-(Ix*fx*(Y-c2))/Z;
    J[3] = (Ix*fx)/(Z*s);
    J[4] = (Iy*fy)/(Z*s);
    J[5] = -(1.0f/(Z*Z)*(Ix*X*fx+Iy*Y*fy))/s;
  }

  inline __m128 div_ps(__m128 a, __m128 b_) const {
    return use_rcp ? _mm_mul_ps(a, _mm_rcp_ps(b_)) : _mm_div_ps(a, b_);   // rigid_body_warp.cc:47-58
  }

  // computeJacobian (rigid_body_warp.cc:60-315): same SSE op sequence per column, 4 points at a time,
  // written straight into the row-major N x 6 output (the reference transposes a column-major temp).
  int compute_jacobian(const float* points, int N, const float* IxIy, float* ret) const {
    float fx = K(0,0), fy = K(1,1);
    float s = T(0,0), c1 = T_inv(0,3), c2 = T_inv(1,3), c3 = T_inv(2,3);
    const __m128 FX = _mm_set1_ps(fx), FY = _mm_set1_ps(fy), C1 = _mm_set1_ps(c1), C2 = _mm_set1_ps(c2),
                 C3 = _mm_set1_ps(c3), S = _mm_set1_ps(s), SIGN = _mm_set1_ps(-0.0f),
                 s_i = _mm_set1_ps((float) (1.0 / (double) s));
    int i = 0;
    for (; i <= N - 4; i += 4) {
      const float* q = points + 4 * (size_t) i; const float* g = IxIy + 2 * (size_t) i;
      __m128 x = _mm_setr_ps(q[0], q[4], q[8], q[12]);
      __m128 y = _mm_setr_ps(q[1], q[5], q[9], q[13]);
      __m128 z = _mm_setr_ps(q[2], q[6], q[10], q[14]);
      __m128 Ix = _mm_mul_ps(FX, _mm_setr_ps(g[0], g[2], g[4], g[6]));
      __m128 Iy = _mm_mul_ps(FY, _mm_setr_ps(g[1], g[3], g[5], g[7]));
      __m128 xIx_yIy = _mm_add_ps(_mm_mul_ps(x, Ix), _mm_mul_ps(y, Iy));
      __m128 z2 = _mm_mul_ps(z, z);
      // column 0 (:66-108)
      __m128 a0 = div_ps(_mm_mul_ps(xIx_yIy, _mm_sub_ps(y, C2)), z2);
      __m128 t2 = div_ps(_mm_mul_ps(Iy, _mm_sub_ps(z, C3)), z);
      __m128 j0 = _mm_sub_ps(_mm_xor_ps(t2, SIGN), a0);
      // column 1 (:111-151)
      __m128 t0 = div_ps(_mm_mul_ps(Ix, _mm_sub_ps(z, C3)), z);
      __m128 t3 = div_ps(_mm_mul_ps(xIx_yIy, _mm_sub_ps(x, C1)), z2);
      __m128 j1 = _mm_add_ps(t0, t3);
      // column 2 (:153-193)
      __m128 j2 = div_ps(_mm_sub_ps(_mm_mul_ps(Iy, _mm_sub_ps(x, C1)), _mm_mul_ps(Ix, _mm_sub_ps(y, C2))), z);
      // columns 3, 4 (:195-248)
      __m128 zs = _mm_mul_ps(z, S);
      __m128 j3 = div_ps(Ix, zs), j4 = div_ps(Iy, zs);
      // column 5 (:250-298)
      __m128 j5 = _mm_xor_ps(div_ps(_mm_mul_ps(s_i, xIx_yIy), z2), SIGN);
      alignas(16) float c[6][4];
      _mm_store_ps(c[0], j0); _mm_store_ps(c[1], j1); _mm_store_ps(c[2], j2);
      _mm_store_ps(c[3], j3); _mm_store_ps(c[4], j4); _mm_store_ps(c[5], j5);
      for (int l = 0; l < 4; ++l) for (int k = 0; k < 6; ++k) ret[6 * (size_t) (i + l) + k] = c[k][l];
    }
    return N;   // "this now is a multiple of 16 always" (:310-313)
  }
};

// HartlyNormalization (bpvo/warps.cc:27-48): sequential float sums over 4-vectors
static M44 hartley_normalization(const float* pts, size_t n) {
  float c[4] = {0, 0, 0, 0};
  for (size_t i = 0; i < n; ++i) for (int k = 0; k < 4; ++k) c[k] += pts[4*i + k];
  for (int k = 0; k < 4; ++k) c[k] /= (float) n;
  float m = 0.0f;
  for (size_t i = 0; i < n; ++i) {
    float d0 = pts[4*i] - c[0], d1 = pts[4*i+1] - c[1], d2 = pts[4*i+2] - c[2], d3 = pts[4*i+3] - c[3];
    m += std::sqrt(d0*d0 + d1*d1 + d2*d2 + d3*d3);
  }
  m /= (float) n;
  float s = (float) (std::sqrt(3.0) / (double) std::max(m, 1e-6f));
  M44 r; std::fill_n(r.m, 16, 0.0f);
  r(0,0) = r(1,1) = r(2,2) = s;
  r(0,3) = -s * c[0]; r(1,3) = -s * c[1]; r(2,3) = -s * c[2];
  r(3,3) = 1.0f;
  return r;
}

// ------------------------------------------------------------------------------------------
// descriptors
// ------------------------------------------------------------------------------------------
struct Descriptor {
  int rows = 0, cols = 0, channels = 0;
  avec<float> planes;   // planar channels x rows x cols
  const float* channel(int c) const { return planes.data() + (size_t) c * rows * cols; }
};

// cv::GaussianBlur(src, dst, Size(3,3), s, s) on CV_8U (census.cc:64-65) -- THIRD-PARTY arithmetic, restated from
// OpenCV 4.x's fixed-point path and pinned bit-exact against cv2 4.13.0 golden vectors (tests/golden/cv2_golden.npz):
// taps k = getGaussianKernel(3, s) = [a, 1 - 2a', a] with a = e / (1 + 2e), e = exp(-1 / (2 s^2)), quantised to 8
// fractional bits (a8 = round(256 a), centre = 256 - 2 a8 so that the taps sum to exactly 1); separable, row pass kept
// in 8.8 fixed point, column pass in 16.16, one rounding at the end: (v + 2^15) >> 16; BORDER_REFLECT_101.
// (OpenCV 2.4.11, the version the reference's README names, filters u8 through a different fixed-point path; the
// difference, if any, is a rounding unit in some pixels -- "parity unpinned by the reference" for sigma_ct > 0.)
static void gaussian_blur3_u8(const uint8_t* src, int rows, int cols, double sigma, uint8_t* dst) {
  const double e = std::exp(-0.5 / (sigma * sigma));
  const int a8 = (int) std::lrint(256.0 * (e / (1.0 + 2.0 * e))), c8 = 256 - 2 * a8;
  auto refl = [](int i, int n) { if (n == 1) return 0; while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * (n - 1) - i; } return i; };
  std::vector<int> h((size_t) rows * cols);
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      const uint8_t* s = src + (size_t) y * cols;
      h[(size_t) y * cols + x] = c8 * (int) s[x] + a8 * ((int) s[refl(x - 1, cols)] + (int) s[refl(x + 1, cols)]);
    }
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      const int v = c8 * h[(size_t) y * cols + x] + a8 * (h[(size_t) refl(y - 1, rows) * cols + x] + h[(size_t) refl(y + 1, rows) * cols + x]);
      const int q = (v + (1 << 15)) >> 16;
      dst[(size_t) y * cols + x] = (uint8_t) (q > 255 ? 255 : q);
    }
}

static void compute_descriptor(const orc_params& p, const uint8_t* img, int rows, int cols, Descriptor& d) {
  d.rows = rows; d.cols = cols;
  const size_t n = (size_t) rows * cols;
  if (p.descriptor == ORC_INTENSITY) {              // intensity_descriptor.cc:31-43 (convertTo CV_32F)
    d.channels = 1; d.planes.resize(n);
    for (size_t i = 0; i < n; ++i) d.planes[i] = (float) img[i];
  } else if (p.descriptor == ORC_BITPLANES) {       // bitplanes_descriptor.cc:84-91
    d.channels = 8; d.planes.resize(8 * n);
    std::vector<uint8_t> C(n), blurred;
    if (p.sigmaPriorToCensusTransform > 0.0f) {       // census.cc:63-65
      blurred.resize(n); gaussian_blur3_u8(img, rows, cols, (double) p.sigmaPriorToCensusTransform, blurred.data()); img = blurred.data();
    }
    census(img, rows, cols, C.data());
    const int nt = std::max(1, p.num_threads);
    (void) nt;
#if defined(_OPENMP)
#pragma omp parallel for num_threads(nt) if (nt > 1)
#endif
    for (int b = 0; b < 8; ++b) {                   // parallel_for(Range(0,8)) bitplanes_descriptor.cc:89-90
      float* dst = d.planes.data() + (size_t) b * n;
      for (size_t i = 0; i < n; ++i) dst[i] = 1.0f * (float) ((C[i] & (1 << b)) >> b) - 0.0f;   // ExtractChannel :37-57
      if (p.sigmaBitPlanes > 0.0f) {
        std::vector<float> tmp(dst, dst + n);
        gaussian_blur5(tmp.data(), rows, cols, p.sigmaBitPlanes, dst);
      }
    }
  } else if (p.descriptor == ORC_INTENSITY_AND_GRADIENT) {     // GradientDescriptor::compute (gradient_descriptor.cc:42-64), sigma = sigmaPriorToCensusTransform (dense_descriptor.cc:49)
    d.channels = 3; d.planes.resize(3 * n);
    float* I0 = d.planes.data();
    for (size_t i = 0; i < n; ++i) I0[i] = (float) img[i];
    std::vector<float> sm;
    const float* I = I0;                                       // the intensity channel stays unsmoothed
    if (p.sigmaPriorToCensusTransform > 0.0f) { sm.resize(n); gaussian_blur_f32(I0, rows, cols, 0, (double) p.sigmaPriorToCensusTransform, sm.data()); I = sm.data(); }
    xgradient(I, rows, cols, I0 + n);
    ygradient(I, rows, cols, I0 + 2 * n);
  } else if (p.descriptor == ORC_DESCRIPTOR_FIELDS) {          // DescriptorFields::compute (gradient_descriptor.cc:101-116), sigmas dfSigma1 / dfSigma2 (types.cc:36-37)
    d.channels = 5; d.planes.resize(5 * n);
    float* I0 = d.planes.data();
    for (size_t i = 0; i < n; ++i) I0[i] = (float) img[i];
    std::vector<float> sm, buf(n), tmp(n);
    const float* I = I0;
    if (p.dfSigma1 > 0.0f) { sm.resize(n); imsmooth(I0, rows, cols, (double) p.dfSigma1, sm.data()); I = sm.data(); }
    for (int g = 0; g < 2; ++g) {
      if (g == 0) xgradient(I, rows, cols, buf.data()); else ygradient(I, rows, cols, buf.data());
      float* pos = I0 + (size_t) (1 + 2 * g) * n; float* neg = I0 + (size_t) (2 + 2 * g) * n;       // splitPosNeg :78-99
      for (size_t i = 0; i < n; ++i) { pos[i] = buf[i] >= 0 ? buf[i] : 0.0f; neg[i] = buf[i] < 0 ? buf[i] : 0.0f; }
      if (p.dfSigma2 > 0.0f) {
        tmp.assign(pos, pos + n); imsmooth(tmp.data(), rows, cols, (double) p.dfSigma2, pos);
        tmp.assign(neg, neg + n); imsmooth(tmp.data(), rows, cols, (double) p.dfSigma2, neg);
      }
    }
  } else {
    throw std::logic_error("oracle: descriptor type not on the hot path");
  }
}

// ------------------------------------------------------------------------------------------
// TemplateData (bpvo/template_data.cc:37-189) + PhotoError standard version (photo_error.cc:336-459)
// ------------------------------------------------------------------------------------------
static inline int Floor(double v) { int i = static_cast<int>(v); return i - (i > v); }   // photo_error.cc:255-265

struct TemplateData {
  int level = 0; orc_params params; Warp warp;
  avec<float> points;     // N x 4
  avec<float> pixels;     // C*N
  avec<float> jacobians;  // (C*N+1) x 6
  std::vector<int32_t> inds;
  avec<float> saliency;
  int N = 0, C = 0;

  void set_data(const Descriptor& desc, const float* D, int Dcols) {
    const int rows = desc.rows, cols = desc.cols;
    saliency.assign((size_t) rows * cols, 0.0f);
    saliency_map(desc.planes.data(), desc.channels, rows, cols, saliency.data());

    IsLocalMax is_local_max{nullptr, cols, -1};
    if (rows * cols >= params.minNumPixelsForNonMaximaSuppression) {
      is_local_max.ptr = saliency.data(); is_local_max.stride = cols; is_local_max.radius = params.nonMaxSuppRadius;
    }
    const int border = std::max(params.nonMaxSuppRadius, 3);
    std::vector<uint16_t> sel;
    for (int y = border; y < rows - border - 1; ++y) {
      const float* srow = saliency.data() + (size_t) y * cols;
      for (int x = border; x < cols - border - 1; ++x)
        if (srow[x] >= params.minSaliency && is_local_max(y, x)) { sel.push_back((uint16_t) y); sel.push_back((uint16_t) x); }
    }
    points.resize(0); inds.resize(0);
    for (size_t i = 0; i < sel.size(); i += 2) {
      int y = sel[i], x = sel[i+1];
      float d = D[(size_t) (1 << level) * ((size_t) y * Dcols + x)];
      if (d >= params.minValidDisparity && d <= params.maxValidDisparity) {
        float pt[4]; warp.make_point((float) x, (float) y, d, pt);
        points.insert(points.end(), pt, pt + 4);
        inds.push_back(y * cols + x);
      }
    }
    int extra = (int) (points.size() / 4) % 16;
    if (extra) { points.resize(points.size() - 4 * (size_t) extra); inds.resize(inds.size() - extra); }
    N = (int) (points.size() / 4); C = desc.channels;
    if (params.withNormalization && N > 0) warp.set_normalization(hartley_normalization(points.data(), N));

    pixels.assign((size_t) C * N, 0.0f);
    jacobians.assign(((size_t) C * N + 1) * 6, 0.0f);   // trailing zero Jacobian (:139-141)
    const float NN = 1.0f / 18.0f;
    avec<float> IxIy((size_t) 2 * N + 8);
    for (int c = 0; c < C; ++c) {
      const float* c_ptr = desc.channel(c);
      float* P_ptr = pixels.data() + (size_t) c * N;
      for (int i = 0; i < N; ++i) {
        const int ii = inds[i];
        P_ptr[i] = c_ptr[ii];
        const float* cc = c_ptr + ii;
        if (params.gradientEstimation == ORC_CD3) {
          IxIy[2*i+0] = 0.5f * (cc[1] - cc[-1]);
          IxIy[2*i+1] = 0.5f * (cc[cols] - cc[-cols]);
        } else {
          IxIy[2*i+0] = NN * (1.0f*cc[-2] - 8.0f*cc[-1] + 8.0f*cc[1] - 1.0f*cc[2]);
          IxIy[2*i+1] = NN * (1.0f*cc[-2*cols] - 8.0f*cc[-1*cols] + 8.0f*cc[+1*cols] - 1.0f*cc[2*cols]);
        }
      }
      float* J_ptr = jacobians.data() + (size_t) c * N * 6;
      int i = warp.compute_jacobian(points.data(), N, IxIy.data(), J_ptr);
      for (; i < N; ++i) warp.jacobian(points.data() + 4*(size_t) i, IxIy[2*i], IxIy[2*i+1], J_ptr + 6*(size_t) i);
    }
  }
};

template <typename T> static inline void interp_cubic(T x, T* c) {          // photo_error.cc:267-279
  const T A = T(-0.5);
  c[0] = ((A*(x + 1) - 5*A)*(x + 1) + 8*A)*(x + 1) - 4*A;
  c[1] = ((A + 2)*x - (A + 3))*x*x + 1;
  c[2] = ((A + 2)*(1 - x) - (A + 3))*(1 - x)*(1 - x) + 1;
  c[3] = T(1) - c[0] - c[1] - c[2];
}
template <typename T> static inline void interp_cosine(T x, T* c) {         // photo_error.cc:281-290
  auto m = (T(1) - std::cos(x * M_PI)) / 2.0;
  c[0] = (T) (T(1) - m); c[1] = (T) m;
}
template <typename T> static inline T interp_hermite(const T* y, T mu) {    // photo_error.cc:311-334 (bias=tension=0)
  auto mu2 = mu*mu; auto mu3 = mu*mu2;
  T m0 = (T) (((y[1] - y[0]) * (1 + T(0)) * (1 - T(0)) / 2.0) + ((y[2] - y[1]) * (1 - T(0)) * (1 - T(0)) / 2.0));
  T m1 = (T) (((y[2] - y[1]) * (1 + T(0)) * (1 - T(0)) / 2.0) + ((y[3] - y[2]) * (1 - T(0)) * (1 - T(0)) / 2.0));
  T a0 = 2*mu3 - 3*mu2 + 1, a1 = mu3 - 2*mu2 + mu, a2 = mu3 - mu2, a3 = -2*mu3 + 3*mu2;
  return a0*y[1] + a1*m0 + a2*m1 + a3*y[2];
}

// TemplateData::computeResiduals (template_data.cc:174-189) + PhotoError::Impl init/run
static void compute_residuals(const TemplateData& td, const Descriptor& desc, const M44& pose,
                              avec<float>& residuals, avec<uint16_t>& valid, std::vector<double>& xy, int nt) {
  if (td.N == 0) throw std::logic_error("you should call setData before calling computeResiduals");
  Warp& warp = const_cast<Warp&>(td.warp);      // `mutable RigidBodyWarp _warp` (template_data.h:83)
  warp.set_pose(pose);
  const int N = td.N, C = td.C, rows = desc.rows, cols = desc.cols, interp = td.params.interp;
  valid.resize(N); residuals.resize((size_t) C * N); xy.resize(2 * (size_t) N);
  // init (photo_error.cc:344-363)
  const int border_lo = (interp == ORC_LINEAR || interp == ORC_COSINE) ? 0 : 1;
  const int border_hi = (interp == ORC_LINEAR || interp == ORC_COSINE) ? 1 : 3;
  double P[3][4];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) P[i][j] = (double) warp.P(i,j);
  for (int i = 0; i < N; ++i) {
    const float* X = td.points.data() + 4 * (size_t) i;
    double h[3];
    for (int r = 0; r < 3; ++r) {
      double s = P[r][0] * (double) X[0]; s += P[r][1] * (double) X[1]; s += P[r][2] * (double) X[2]; s += P[r][3] * (double) X[3];
      h[r] = s;
    }
    double w = 1.0 / h[2];                       // normHomog (eigen.h): (1/p[2]) * p.head<2>()
    double u = w * h[0], v = w * h[1];
    xy[2*i] = u; xy[2*i+1] = v;
    int xi = Floor(u), yi = Floor(v);
    valid[i] = xi >= border_lo && xi < cols - border_hi && yi >= border_lo && yi < rows - 1;
  }
  (void) nt;
  // run, per channel (photo_error.cc:365-451; parallel_for over channels template_data.cc:188)
#if defined(_OPENMP)
#pragma omp parallel for num_threads(nt) if (nt > 1)
#endif
  for (int c = 0; c < C; ++c) {
    const float* I0 = td.pixels.data() + (size_t) c * N;
    const float* I1 = desc.channel(c);
    float* r = residuals.data() + (size_t) c * N;
    const int stride = cols;
    for (int i = 0; i < N; ++i) {
      if (!valid[i]) { r[i] = 0.0f; continue; }
      double xf = xy[2*i], yf = xy[2*i+1];
      int xi = Floor(xf), yi = Floor(yf);
      xf -= (double) xi; yf -= (double) yi;
      switch (interp) {
        case ORC_LINEAR: {
          int ii = yi * stride + xi;
          double wx = (1.0 - xf);
          double Iw = (1.0 - yf) * (I1[ii] * wx + I1[ii+1] * xf) + yf * (I1[ii+stride] * wx + I1[ii+stride+1] * xf);
          r[i] = float(Iw - (double) I0[i]);
        } break;
        case ORC_COSINE: {
          float Cx[2], Cy[2];
          const float* p1 = I1 + (yi + 0) * stride + xi; const float* p2 = I1 + (yi + 1) * stride + xi;
          interp_cosine((float) xf, Cx); interp_cosine((float) yf, Cy);
          float d1 = p1[0]*Cx[0] + p1[1]*Cx[1], d2 = p2[0]*Cx[0] + p2[1]*Cx[1];
          float Iw = Cy[0]*d1 + Cy[1]*d2;
          r[i] = Iw - I0[i];
        } break;
        case ORC_CUBIC: {
          float Cx[4], Cy[4], d[4];
          interp_cubic((float) xf, Cx); interp_cubic((float) yf, Cy);
          for (int k = 0; k < 4; ++k) {
            const float* p = I1 + (yi - 1 + k) * stride + xi;
            // NB: the reference maps 4 floats starting AT xi (not xi-1) (photo_error.cc:413-416)
            d[k] = ((p[0]*Cx[0] + p[1]*Cx[1]) + p[2]*Cx[2]) + p[3]*Cx[3];
          }
          float Iw = ((Cy[0]*d[0] + Cy[1]*d[1]) + Cy[2]*d[2]) + Cy[3]*d[3];
          r[i] = Iw - I0[i];
        } break;
        default: {
          float V[4];
          for (int k = 0; k < 4; ++k) V[k] = interp_hermite(I1 + (yi - 1 + k) * stride + xi, (float) xf);
          float Iw = interp_hermite(V, (float) yf);
          r[i] = Iw - I0[i];
        } break;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// scale, weights, normal equations
// ------------------------------------------------------------------------------------------
// median (bpvo/utils.h:224-252)
static float median_of(float* first, size_t n) {
  if (n == 0) return 0.0f;
  if (n < 3) return first[0];
  float* middle = first + n / 2;
  std::nth_element(first, middle, first + n);
  if (n % 2 != 0) return *middle;
  float* m = std::max_element(first, middle);
  return (float) ((*m + *middle) / 2.0);
}

// AutoScaleEstimator (bpvo/mestimator.cc:416-490, mestimator.h:62-83)
struct ScaleEstimator {
  float scale = 1.0f, delta_scale = 1e10f, tol = 1e-6f;
  avec<float> buffer;
  void reset() { delta_scale = 1e10f; scale = 1.0f; }
  float estimate(const avec<float>& r, const avec<uint16_t>& valid) {
    if (delta_scale > tol) {
      buffer.resize(0); buffer.reserve(r.size());
      for (size_t i = 0; i < r.size(); ++i) if (valid[i] != 0) buffer.push_back(std::fabs(r[i]));
      float med = median_of(buffer.data(), buffer.size());
      float s = (1.4826f * (1.0f + 5.0f / (buffer.size() - 6))) * med;     // size_t arithmetic as the reference (:463)
      if (s < 1e-6) s = 1.0f;
      delta_scale = std::fabs(s - scale);
      scale = s;
    }
    return scale;
  }
};

// MEstimator::ComputeWeights, WITH_SIMD + __AVX__ path (bpvo/mestimator.cc:242-415); `valid` ignored (Q6)
static void compute_weights(int loss, const avec<float>& r, const avec<uint16_t>& valid, float sigma, avec<float>& w) {
  w.resize(valid.size());
  if (loss == ORC_L2) { std::fill(w.begin(), w.end(), 1.0f); return; }
  const size_t N = r.size(), n = N & ~(size_t) 15;
  const float sigma_inv = 1.0f / sigma;
  const __m256 s_inv = _mm256_set1_ps(sigma_inv), SIGN = _mm256_set1_ps(-0.f);
  size_t i = 0;
  if (loss == ORC_HUBER) {
    const float huber_k = 1.345f; const __m256 h_k = _mm256_set1_ps(huber_k);
    for (; i < n; i += 8) {
      __m256 x = _mm256_andnot_ps(SIGN, _mm256_mul_ps(_mm256_load_ps(&r[i]), s_inv));
      _mm256_store_ps(&w[i], _mm256_div_ps(h_k, _mm256_max_ps(x, h_k)));
    }
    for (; i < N; ++i) {   // HuberOp scalar tail (:35-47, :294-300)
      float x = std::fabs(sigma_inv * r[i]);
      w[i] = (float) valid[i] * ((x < huber_k) ? 1.0f : (huber_k / x));
    }
  } else if (loss == ORC_TUKEY) {
    const float tukey_t = 4.685f;
    const __m256 ones = _mm256_set1_ps(1.0f), t = _mm256_set1_ps(tukey_t), t_i = _mm256_set1_ps((float) (1.0 / tukey_t));
    for (; i < n; i += 8) {
      __m256 x = _mm256_mul_ps(_mm256_load_ps(&r[i]), s_inv);
      __m256 rr = _mm256_mul_ps(x, t_i);
      rr = _mm256_sub_ps(ones, _mm256_mul_ps(rr, rr));
      rr = _mm256_mul_ps(rr, rr);
      __m256 m = _mm256_cmp_ps(_mm256_andnot_ps(SIGN, x), t, _CMP_LT_OQ);
      _mm256_store_ps(&w[i], _mm256_and_ps(m, rr));
    }
    const float t_inv = (float) (1.0 / tukey_t);
    for (; i < N; ++i) {   // TukeyOp scalar tail (:49-61, :378-384)
      float x = std::fabs(sigma_inv * r[i]);
      float f = (x < 1e-6) ? 1.0f : (x > tukey_t) ? 0.0f : ((1.0f - (t_inv*x)*(t_inv*x)) * (1.0f - (t_inv*x)*(t_inv*x)));
      w[i] = (float) valid[i] * f;
    }
  } else {
    throw std::logic_error("unknown RobustFunction");
  }
  _mm256_zeroupper();
}

// LinearSystemBuilderReduction::rankUpdatePoint (bpvo/linear_system_builder.cc:140-205), SSE
static inline void rank_update(const float* J, float Ri, float Wi, uint16_t Vi, float* data, float* G, float& res_norm) {
  float w = Wi * static_cast<float>(Vi);
  float wR = w * Ri;
  __m128 wwww = _mm_set1_ps(w);
  __m128 v1234 = _mm_loadu_ps(J);
  __m128 v56xx = _mm_loadu_ps(J + 4);
  __m128 v1212 = _mm_movelh_ps(v1234, v1234);
  __m128 v3434 = _mm_movehl_ps(v1234, v1234);
  __m128 v5656 = _mm_movelh_ps(v56xx, v56xx);
  __m128 v1122 = _mm_mul_ps(wwww, _mm_unpacklo_ps(v1212, v1212));
  _mm_store_ps(data + 0, _mm_add_ps(_mm_load_ps(data + 0), _mm_mul_ps(v1122, v1212)));
  _mm_store_ps(data + 4, _mm_add_ps(_mm_load_ps(data + 4), _mm_mul_ps(v1122, v3434)));
  _mm_store_ps(data + 8, _mm_add_ps(_mm_load_ps(data + 8), _mm_mul_ps(v1122, v5656)));
  __m128 v3344 = _mm_mul_ps(wwww, _mm_unpacklo_ps(v3434, v3434));
  _mm_store_ps(data + 12, _mm_add_ps(_mm_load_ps(data + 12), _mm_mul_ps(v3344, v3434)));
  _mm_store_ps(data + 16, _mm_add_ps(_mm_load_ps(data + 16), _mm_mul_ps(v3344, v5656)));
  __m128 v5566 = _mm_mul_ps(wwww, _mm_unpacklo_ps(v5656, v5656));
  _mm_store_ps(data + 20, _mm_add_ps(_mm_load_ps(data + 20), _mm_mul_ps(v5566, v5656)));
  __m128 g1 = _mm_load_ps(G), g2 = _mm_load_ps(G + 4);
  __m128 wr = _mm_mul_ps(wwww, _mm_set1_ps(Ri));
  _mm_store_ps(G, _mm_add_ps(g1, _mm_mul_ps(wr, v1234)));
  _mm_store_ps(G + 4, _mm_add_ps(g2, _mm_mul_ps(wr, v56xx)));
  res_norm += wR * Ri;
}

// toEigen (linear_system_builder.cc:207-221): unpack the 24-float upper 2x2 blocks; column-major out
static void unpack_hessian(const float* data, float* H) {
  float R[6][6] = {{0}};
  for (int i = 0, ii = 0; i < 6; i += 2)
    for (int j = i; j < 6; j += 2) {
      R[i][j] = data[ii++]; R[i][j+1] = data[ii++]; R[i+1][j] = data[ii++]; R[i+1][j+1] = data[ii++];
    }
  for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { H[j*6 + i] = R[i][j]; H[i*6 + j] = R[i][j]; }
}

// LinearSystemBuilder::Run (linear_system_builder.cc:334-350, 223-267).  nt == 1: the sequential
// non-TBB path.  nt > 1: range split + join in range order (stand-in for tbb::parallel_reduce :96-130).
static float build_linear_system(const avec<float>& J, const avec<float>& R, const avec<float>& W,
                                 const avec<uint16_t>& V, float* H, float* G, int nt) {
  const size_t n = R.size();
  nt = std::max(1, nt);
  std::vector<float> datas((size_t) nt * 24, 0.0f), gs((size_t) nt * 8, 0.0f), rs(nt, 0.0f);
#if defined(_OPENMP)
#pragma omp parallel for num_threads(nt) if (nt > 1)
#endif
  for (int t = 0; t < nt; ++t) {
    alignas(16) float data[24]; alignas(16) float Gd[8];
    std::fill_n(data, 24, 0.0f); std::fill_n(Gd, 8, 0.0f);
    float ret = 0.0f;
    size_t lo = n * t / nt, hi = n * (t + 1) / nt;
    for (size_t i = lo; i < hi; ++i) rank_update(J.data() + 6 * i, R[i], W[i], V[i], data, Gd, ret);
    memcpy(&datas[(size_t) t * 24], data, sizeof(data)); memcpy(&gs[(size_t) t * 8], Gd, sizeof(Gd)); rs[t] = ret;
  }
  float Hs[36]; std::fill_n(Hs, 36, 0.0f); std::fill_n(G, 6, 0.0f);
  float ret = 0.0f;
  for (int t = 0; t < nt; ++t) {
    float Ht[36]; unpack_hessian(&datas[(size_t) t * 24], Ht);
    for (int i = 0; i < 36; ++i) Hs[i] += Ht[i];
    for (int i = 0; i < 6; ++i) G[i] += gs[(size_t) t * 8 + i];
    ret += rs[t];
  }
  memcpy(H, Hs, sizeof(Hs));
  return std::sqrt(ret);
}

}  // namespace

// ==========================================================================================
// frame / estimator / vo objects
// ==========================================================================================
struct orc_frame {
  orc_params params; int rows, cols;
  bool has_data = false, has_template = false;
  std::vector<uint8_t> image; std::vector<float> disparity;
  std::vector<std::vector<uint8_t>> pyr; std::vector<int> prow, pcol;
  std::vector<Descriptor> desc;
  std::vector<TemplateData> tdata;

  // VisualOdometryFrame ctor (bpvo/vo_frame.cc:13-30): K halves (K(2,2)=1), baseline doubles per level
  orc_frame(const float K[9], float b, int r, int c, const orc_params& p) : params(p), rows(r), cols(c) {
    const int L = p.numPyramidLevels;
    pyr.resize(L); prow.resize(L); pcol.resize(L); desc.resize(L); tdata.resize(L);
    M33 Kp; memcpy(Kp.m, K, sizeof(Kp.m)); float bp = b;
    int rr = r, cc = c;
    for (int i = 0; i < L; ++i) {
      if (i > 0) { for (int k = 0; k < 9; ++k) Kp.m[k] *= 0.5f; Kp(2,2) = 1.0f; bp *= 2.0f; rr = (rr + 1) / 2; cc = (cc + 1) / 2; }
      prow[i] = rr; pcol[i] = cc;
      tdata[i].level = i; tdata[i].params = p; tdata[i].warp.init(Kp, bp); tdata[i].warp.use_rcp = p.use_rcp != 0;
    }
  }
  // setData (vo_frame.cc:48-55) -> DenseDescriptorPyramid::init (dense_descriptor_pyramid.cc:67-78)
  void set_data(const uint8_t* I, const float* D) {
    image.assign(I, I + (size_t) rows * cols); disparity.assign(D, D + (size_t) rows * cols);
    const int L = params.numPyramidLevels;
    pyr[0] = image;
    for (int i = 1; i < L; ++i) { pyr[i].resize((size_t) prow[i] * pcol[i]); pyr_down(pyr[i-1].data(), prow[i-1], pcol[i-1], pyr[i].data()); }
    for (int i = L - 1; i >= params.maxTestLevel; --i) compute_descriptor(params, pyr[i].data(), prow[i], pcol[i], desc[i]);
    has_data = true;
  }
  // setTemplate (vo_frame.cc:61-93)
  void set_template() {
    if (!has_data) throw std::logic_error("no data in frame");
    for (int i = (int) tdata.size() - 1; i >= params.maxTestLevel; --i) tdata[i].set_data(desc[i], disparity.data(), cols);
    has_template = true;
  }
};

struct orc_estimator {
  orc_params params;
  ScaleEstimator scale;
  avec<float> residuals, weights; avec<uint16_t> valid;
  std::vector<double> xy;
  float f_norm_prev = 0.0f, g_tol = 0.0f; int num_fun_evals = 0; int total_fun_evals = 0;
  float last_sigma = 1.0f;
  std::vector<float> trace; bool trace_on = false;    // rows {level, eval, f_norm, |dp|, max|G|, sigma, converged, status}: what
                                                       // run() prints at verbosity kIteration (pose_estimator_base.h:231-247)
  explicit orc_estimator(const orc_params& p) : params(p) {}

  void reset() { scale.reset(); f_norm_prev = 0.0f; g_tol = 0.0f; num_fun_evals = 0; }   // pose_estimator_base.h:287-293

  // PoseEstimatorGN::linearize (pose_estimator_gn.h:70-81)
  float linearize(const TemplateData& td, const Descriptor& desc, const M44& T, float* H, float* G) {
    compute_residuals(td, desc, T, residuals, valid, xy, params.num_threads);
    if (residuals.size() != valid.size()) {          // replicateValidFlags (pose_estimator_base.h:307-320)
      size_t m = residuals.size() / valid.size(), nv = valid.size();
      avec<uint16_t> tmp(residuals.size());
      for (size_t i = 0; i < m; ++i) memcpy(tmp.data() + i * nv, valid.data(), nv * sizeof(uint16_t));
      valid.swap(tmp);
    }
    float sigma = scale.estimate(residuals, valid);
    last_sigma = sigma;
    compute_weights(params.lossFunction, residuals, valid, sigma, weights);
    num_fun_evals += 1; total_fun_evals += 1;
    return build_linear_system(td.jacobians, residuals, weights, valid, H, G, params.num_threads);
  }

  bool test_convergence(float dp_norm, float dp_norm_prev, float g_norm, float f_norm, int& status) const {   // :258-282
    static const float sqrt_eps = std::sqrt(std::numeric_limits<float>::epsilon());
    if (dp_norm < params.parameterTolerance || dp_norm < params.parameterTolerance * (sqrt_eps + dp_norm_prev)) { status = ORC_PARAM_TOL; return true; }
    if (f_norm < params.functionTolerance || f_norm < params.functionTolerance * (sqrt_eps + f_norm_prev) ||
        std::fabs(f_norm - f_norm_prev) < params.functionTolerance) { status = ORC_FUNC_TOL; return true; }
    if (g_norm < g_tol) { status = ORC_GRAD_TOL; return true; }
    return false;
  }

  // PoseEstimatorBase::run (pose_estimator_base.h:324-407)
  orc_stats run(const TemplateData& td, const Descriptor& desc, M44& T) {
    reset();
    orc_stats ret; ret.numIterations = 0; ret.firstOrderOptimality = 0; ret.status = ORC_MAX_ITERS; ret.finalError = -1.0f;
    const int maxFuncEvals = 6 * 200;                 // pose_estimator_params.h:33
    M44 dT = T; float H[36], G[6], dp[6];
    auto gnorm = [&]() { float g = 0; for (int i = 0; i < 6; ++i) g = std::max(g, std::fabs(G[i])); return g; };
    float f_norm = linearize(td, desc, dT, H, G), g_norm = gnorm();
    g_tol = params.gradientTolerance * std::max(g_norm, std::sqrt(std::numeric_limits<float>::epsilon()));
    if (g_norm < g_tol) { ret.status = ORC_GRAD_TOL; ret.finalError = f_norm; ret.numIterations = 1; ret.firstOrderOptimality = g_norm; return ret; }
    if (!solve6(H, G, dp)) { ret.status = ORC_SOLVER_ERROR; ret.finalError = f_norm; return ret; }
    f_norm_prev = 0.0f; float dp_norm_prev = 0.0f; bool has_converged = false;
    auto neg = [&](float* o) { for (int i = 0; i < 6; ++i) o[i] = -dp[i]; };
    float ndp[6]; neg(ndp); dT = mul(dT, td.warp.params_to_pose(ndp));
    do {
      float dp_norm = 0; for (int i = 0; i < 6; ++i) dp_norm += dp[i]*dp[i]; dp_norm = std::sqrt(dp_norm);
      g_norm = gnorm();
      has_converged = test_convergence(dp_norm, dp_norm_prev, g_norm, f_norm, ret.status);
      if (trace_on) { const float row[8] = {(float) td.level, (float) num_fun_evals, f_norm, dp_norm, g_norm, last_sigma, has_converged ? 1.0f : 0.0f, (float) ret.status}; trace.insert(trace.end(), row, row + 8); }
      dp_norm_prev = dp_norm; f_norm_prev = f_norm;
      if (!has_converged) {                            // runIteration (pose_estimator_gn.h:83-100)
        f_norm = linearize(td, desc, dT, H, G);
        if (!solve6(H, G, dp)) { ret.status = ORC_SOLVER_ERROR; break; }
      }
      neg(ndp); dT = mul(dT, td.warp.params_to_pose(ndp));          // also when converged (Q1)
    } while (ret.numIterations++ < params.maxIterations && !has_converged && num_fun_evals < maxFuncEvals);
    if (ret.status != ORC_SOLVER_ERROR) T = dT;
    ret.numIterations -= 1; ret.finalError = f_norm; ret.firstOrderOptimality = g_norm;
    return ret;
  }
};

struct orc_vo {
  orc_params params; int rows, cols;
  std::unique_ptr<orc_estimator> est;
  std::unique_ptr<orc_frame> ref, cur, prev;
  M44 T_kf; std::vector<M44> trajectory;
  std::vector<float> pc_xyzw, pc_w; std::vector<uint8_t> pc_gray;
};

namespace {

// VisualOdometryPoseEstimator::estimatePose (bpvo/vo_pose_estimator.cc:63-93)
static int estimate_pose(orc_estimator& est, const orc_frame& ref, const orc_frame& cur, const M44& T_init, M44& T_est, orc_stats* stats) {
  const int L = ref.params.numPyramidLevels;
  for (int i = 0; i < L; ++i) { stats[i].numIterations = 0; stats[i].finalError = -1.0f; stats[i].firstOrderOptimality = -1.0f; stats[i].status = ORC_SOLVER_ERROR; }
  T_est = T_init;
  int evals0 = est.total_fun_evals;
  for (int i = L - 1; i >= est.params.maxTestLevel; --i) stats[i] = est.run(ref.tdata[i], cur.desc[i], T_est);
  return est.total_fun_evals - evals0;
}

static float fraction_good(const orc_estimator& est, float thresh) {   // vo_pose_estimator.cc:101-107
  size_t n = 0; for (float w : est.weights) n += (w > thresh);
  return n / static_cast<float>(est.weights.size());
}

// Trajectory::push_back with its InvertPose (bpvo/trajectory.cc:30-50) -- NB translation is -R*t (sic)
static void trajectory_push(std::vector<M44>& poses, const M44& T) {
  M44 inv; std::fill_n(inv.m, 16, 0.0f);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) inv(i,j) = T(j,i);
  for (int i = 0; i < 3; ++i) { float s = T(i,0)*T(0,3); s += T(i,1)*T(1,3); s += T(i,2)*T(2,3); inv(i,3) = -s; }
  inv(3,3) = 1.0f;
  if (!poses.empty()) poses.push_back(mul(poses.back(), inv)); else poses.push_back(inv);
}

// shouldKeyFrame (bpvo/vo.cc:199-224), RotationMatrixToEulerAngles (math_utils.h:210-222)
static int should_key_frame(const orc_vo& vo, const M44& pose) {
  const orc_params& p = vo.params;
  float t_norm = pose(0,3)*pose(0,3) + pose(1,3)*pose(1,3) + pose(2,3)*pose(2,3);
  if (t_norm > p.minTranslationMagToKeyFrame * p.minTranslationMagToKeyFrame) return ORC_KF_LARGE_TRANSLATION;
  float eta = (float) (1.0 / (std::sqrt(pose(0,0)*pose(0,0) + pose(1,0)*pose(1,0))));
  float rz = std::asin(eta * pose(1,0)), ry = std::asin(-pose(2,0)), rx = std::asin(eta * pose(2,1));
  float r_norm = rx*rx + ry*ry + rz*rz;
  if (r_norm > p.minRotationMagToKeyFrame * p.minRotationMagToKeyFrame) return ORC_KF_LARGE_ROTATION;   // radians vs degrees (Q8)
  float frac_good = fraction_good(*vo.est, p.goodPointThreshold);
  if (frac_good < p.maxFractionOfGoodPointsToKeyFrame) return ORC_KF_SMALL_FRAC_GOOD;
  return ORC_KF_NONE;
}

// getPointCloudFromRefFrame (bpvo/vo.cc:243-281)
static void point_cloud_from_ref(orc_vo& vo) {
  const TemplateData& td = vo.ref->tdata[vo.params.maxTestLevel];
  const size_t n = td.N;
  if (n > vo.est->weights.size()) throw std::logic_error("size mismatch");
  vo.pc_xyzw.assign(td.points.begin(), td.points.end()); vo.pc_w.resize(n); vo.pc_gray.resize(n);
  const M33& K = td.warp.K;
  for (size_t i = 0; i < n; ++i) {
    const float* X = td.points.data() + 4 * i;
    float x[3];
    for (int r = 0; r < 3; ++r) { float s = K(r,0)*X[0]; s += K(r,1)*X[1]; s += K(r,2)*X[2]; x[r] = s; }
    float z_i = 1.0f / x[2]; float u = z_i * x[0], v = z_i * x[1];
    uint8_t c = (v >= 0 && v < vo.rows && u >= 0 && u < vo.cols) ? vo.ref->image[(size_t) ((int) v) * vo.cols + (int) u] : 0;
    vo.pc_gray[i] = c; vo.pc_w[i] = vo.est->weights[i];
  }
}

}  // namespace

// ==========================================================================================
// C API
// ==========================================================================================
#define ORC_TRY try {
#define ORC_CATCH(rv) } catch (const std::exception& e) { g_last_error = e.what(); return rv; }

extern "C" {

const char* orc_last_error(void) { return g_last_error.c_str(); }

void orc_default_params(orc_params* p) {        // bpvo/types.cc:31-66
  p->numPyramidLevels = -1; p->minImageDimensionForPyramid = 40;
  p->sigmaPriorToCensusTransform = -1.0f; p->sigmaBitPlanes = 0.5f;
  p->maxIterations = 50; p->parameterTolerance = 1e-7f; p->functionTolerance = 1e-6f; p->gradientTolerance = 1e-8f;
  p->relaxTolerancesForCoarseLevels = 1; p->gradientEstimation = ORC_CD3; p->interp = ORC_LINEAR;
  p->lossFunction = ORC_TUKEY; p->descriptor = ORC_INTENSITY; p->verbosity = 0x20;
  p->minTranslationMagToKeyFrame = 0.15f; p->minRotationMagToKeyFrame = 5.0f;
  p->maxFractionOfGoodPointsToKeyFrame = 0.6f; p->goodPointThreshold = 0.85f;
  p->minNumPixelsForNonMaximaSuppression = 320 * 240; p->nonMaxSuppRadius = 1; p->minNumPixelsToWork = 256;
  p->minSaliency = 0.1f; p->minValidDisparity = 0.001f; p->maxValidDisparity = 512.0f;
  p->maxTestLevel = 0; p->withNormalization = 1;
  p->use_rcp = 1; p->num_threads = 1; p->dfSigma1 = 0.75f; p->dfSigma2 = 1.75f;
}

void orc_pyr_down(const uint8_t* src, int rows, int cols, uint8_t* dst) { pyr_down(src, rows, cols, dst); }
void orc_gaussian_blur5(const float* src, int rows, int cols, float sigma, float* dst) { gaussian_blur5(src, rows, cols, sigma, dst); }
void orc_gaussian_blur_f32(const float* src, int rows, int cols, int ksize, double sigma, float* dst) { gaussian_blur_f32(src, rows, cols, ksize, sigma, dst); }
void orc_census(const uint8_t* src, int rows, int cols, uint8_t* dst) { census(src, rows, cols, dst); }
void orc_gaussian_blur3_u8(const uint8_t* src, int rows, int cols, float sigma, uint8_t* dst) { gaussian_blur3_u8(src, rows, cols, (double) sigma, dst); }
int orc_descriptor(const orc_params* p, const uint8_t* img, int rows, int cols, float* planes) {
  ORC_TRY
  Descriptor d; compute_descriptor(*p, img, rows, cols, d);
  memcpy(planes, d.planes.data(), d.planes.size() * sizeof(float));
  return d.channels;
  ORC_CATCH(-1)
}
void orc_saliency(const float* planes, int channels, int rows, int cols, float* dst) { saliency_map(planes, channels, rows, cols, dst); }
float orc_median(float* buf, size_t n) { return median_of(buf, n); }
int orc_solve6(const float H[36], const float G[6], float dp[6]) { return solve6(H, G, dp) ? 1 : 0; }
void orc_params_to_pose(const float Tn[16], const float p[6], float out[16]) {
  Warp w; M44 T; memcpy(T.m, Tn, sizeof(T.m)); w.set_normalization(T);
  M44 r = w.params_to_pose(p); memcpy(out, r.m, sizeof(r.m));
}

orc_frame* orc_frame_create(const float K[9], float baseline, int rows, int cols, const orc_params* p) {
  ORC_TRY
  if (p->numPyramidLevels <= 0) throw std::logic_error("invalid number of pyramid levels");   // dense_descriptor_pyramid.cc:37
  if (p->maxTestLevel < 0) throw std::logic_error("invalid maxTestLevel");
  return new orc_frame(K, baseline, rows, cols, *p);
  ORC_CATCH(nullptr)
}
void orc_frame_destroy(orc_frame* f) { delete f; }
void orc_frame_set_data(orc_frame* f, const uint8_t* image, const float* disparity) { f->set_data(image, disparity); }
int orc_frame_set_template(orc_frame* f) { ORC_TRY f->set_template(); return 0; ORC_CATCH(-1) }
int orc_frame_num_levels(const orc_frame* f) { return f->params.numPyramidLevels; }
void orc_frame_level_size(const orc_frame* f, int l, int* rows, int* cols) { *rows = f->prow[l]; *cols = f->pcol[l]; }
const uint8_t* orc_frame_pyramid(const orc_frame* f, int l) { return f->pyr[l].data(); }
const float* orc_frame_descriptor(const orc_frame* f, int l, int* ch) { *ch = f->desc[l].channels; return f->desc[l].planes.data(); }
const float* orc_frame_saliency(const orc_frame* f, int l) { return f->tdata[l].saliency.data(); }
int orc_frame_num_points(const orc_frame* f, int l) { return f->tdata[l].N; }
const float* orc_frame_points(const orc_frame* f, int l) { return f->tdata[l].points.data(); }
const float* orc_frame_pixels(const orc_frame* f, int l) { return f->tdata[l].pixels.data(); }
const float* orc_frame_jacobians(const orc_frame* f, int l) { return f->tdata[l].jacobians.data(); }
const int32_t* orc_frame_point_inds(const orc_frame* f, int l) { return f->tdata[l].inds.data(); }
void orc_frame_normalization(const orc_frame* f, int l, float Tn[16]) { memcpy(Tn, f->tdata[l].warp.T.m, 16 * sizeof(float)); }

orc_estimator* orc_estimator_create(const orc_params* p) { return new orc_estimator(*p); }
void orc_estimator_destroy(orc_estimator* e) { delete e; }
void orc_estimator_set_trace(orc_estimator* e, int on) { e->trace_on = on != 0; e->trace.clear(); }
int orc_estimator_get_trace(orc_estimator* e, float* rows, int max_rows) {
  const int n = (int) (e->trace.size() / 8);
  if (rows) memcpy(rows, e->trace.data(), sizeof(float) * 8 * (size_t) std::min(n, max_rows));
  return n;
}
float orc_linearize(orc_estimator* e, const orc_frame* ref, const orc_frame* cur, int level, const float T[16],
                    int reset_scale, float H[36], float G[6], float* sigma) {
  ORC_TRY
  if (reset_scale) e->reset();
  M44 Tm; memcpy(Tm.m, T, sizeof(Tm.m));
  float f = e->linearize(ref->tdata[level], cur->desc[level], Tm, H, G);
  if (sigma) *sigma = e->last_sigma;
  return f;
  ORC_CATCH(-1.0f)
}
size_t orc_estimator_num_residuals(const orc_estimator* e) { return e->residuals.size(); }
const float* orc_estimator_residuals(const orc_estimator* e) { return e->residuals.data(); }
const float* orc_estimator_weights(const orc_estimator* e) { return e->weights.data(); }
const uint16_t* orc_estimator_valid(const orc_estimator* e) { return e->valid.data(); }
int orc_estimate_pose(orc_estimator* e, const orc_frame* ref, const orc_frame* cur, const float T_init[16], float T_est[16], orc_stats* stats) {
  ORC_TRY
  M44 Ti, Te; memcpy(Ti.m, T_init, sizeof(Ti.m));
  int n = estimate_pose(*e, *ref, *cur, Ti, Te, stats);
  memcpy(T_est, Te.m, sizeof(Te.m));
  return n;
  ORC_CATCH(-1)
}
float orc_fraction_good(const orc_estimator* e, float thresh) { return fraction_good(*e, thresh); }

// VisualOdometry::Impl ctor (bpvo/vo.cc:94-110)
orc_vo* orc_vo_create(const float K[9], float baseline, int rows, int cols, const orc_params* p) {
  ORC_TRY
  std::unique_ptr<orc_vo> vo(new orc_vo);
  vo->params = *p; vo->rows = rows; vo->cols = cols; vo->T_kf = M44::Identity();
  vo->est.reset(new orc_estimator(*p));
  if (vo->params.numPyramidLevels <= 0)
    vo->params.numPyramidLevels = 1 + (int) std::round(std::log2(std::min(rows, cols) / (double) p->minImageDimensionForPyramid));
  for (auto* fp : {&vo->ref, &vo->cur, &vo->prev}) {
    fp->reset(orc_frame_create(K, baseline, rows, cols, &vo->params));
    if (!*fp) throw std::logic_error(g_last_error);
  }
  return vo.release();
  ORC_CATCH(nullptr)
}
void orc_vo_destroy(orc_vo* vo) { delete vo; }

// VisualOdometry::Impl::addFrame (bpvo/vo.cc:125-197)
int orc_vo_add_frame(orc_vo* vo, const uint8_t* image, const float* disparity, orc_result* out) {
  ORC_TRY
  if (image == nullptr || disparity == nullptr) throw std::logic_error("nullptr image/disparity");   // vo.cc:68
  const int L = vo->params.numPyramidLevels;
  memset(out, 0, sizeof(*out)); out->numLevels = L;
  for (int i = 0; i < L && i < 16; ++i) { out->stats[i].finalError = -1.0f; out->stats[i].firstOrderOptimality = -1.0f; out->stats[i].status = ORC_SOLVER_ERROR; }
  M44 I4 = M44::Identity(); memcpy(out->pose, I4.m, sizeof(I4.m));
  vo->pc_xyzw.clear(); vo->pc_w.clear(); vo->pc_gray.clear();
  vo->cur->set_data(image, disparity);
  if (!vo->ref->has_template) {
    std::swap(vo->ref, vo->cur);
    vo->ref->set_template();
    trajectory_push(vo->trajectory, vo->T_kf);
    out->isKeyFrame = 1; out->keyFramingReason = ORC_KF_FIRST_FRAME;
    return 0;
  }
  M44 T_est; std::vector<orc_stats> stats(L);
  out->numFunEvals += estimate_pose(*vo->est, *vo->ref, *vo->cur, vo->T_kf, T_est, stats.data());
  int reason = should_key_frame(*vo, T_est);
  out->keyFramingReason = reason; out->isKeyFrame = reason != ORC_KF_NONE;
  M44 pose;
  if (!out->isKeyFrame) {
    std::swap(vo->prev, vo->cur);
    pose = mul(T_est, inverse(vo->T_kf));
    vo->T_kf = T_est;
  } else {
    point_cloud_from_ref(*vo);
    if (!vo->prev->has_data) {
      std::swap(vo->cur, vo->ref);
      vo->ref->set_template();
      pose = mul(T_est, inverse(vo->T_kf));
      vo->T_kf = M44::Identity();
    } else {
      std::swap(vo->prev, vo->ref);
      vo->prev->has_data = false; vo->prev->has_template = false;     // clear()
      vo->ref->set_template();
      M44 T_init = M44::Identity();
      out->numFunEvals += estimate_pose(*vo->est, *vo->ref, *vo->cur, T_init, T_est, stats.data());
      pose = T_est; vo->T_kf = T_est;
    }
  }
  trajectory_push(vo->trajectory, pose);
  memcpy(out->pose, pose.m, sizeof(pose.m));
  for (int i = 0; i < L && i < 16; ++i) out->stats[i] = stats[i];
  out->numPointCloud = (int) vo->pc_w.size();
  return 0;
  ORC_CATCH(-1)
}
int orc_vo_num_points_at_level(const orc_vo* vo, int level) {
  if (level < 0) level = vo->params.maxTestLevel;
  return vo->ref ? vo->ref->tdata[level].N : 0;
}
const orc_frame* orc_vo_ref_frame(const orc_vo* vo) { return vo->ref.get(); }
const orc_estimator* orc_vo_estimator(const orc_vo* vo) { return vo->est.get(); }
int orc_vo_trajectory(const orc_vo* vo, float* poses, int max_poses) {
  int n = std::min<int>(max_poses, (int) vo->trajectory.size());
  for (int i = 0; i < n; ++i) memcpy(poses + 16 * (size_t) i, vo->trajectory[i].m, 16 * sizeof(float));
  return (int) vo->trajectory.size();
}
int orc_vo_point_cloud(const orc_vo* vo, float* xyzw, float* weights, uint8_t* gray, int max_points) {
  int n = std::min<int>(max_points, (int) vo->pc_w.size());
  if (n > 0) { memcpy(xyzw, vo->pc_xyzw.data(), 16 * (size_t) n); memcpy(weights, vo->pc_w.data(), 4 * (size_t) n); memcpy(gray, vo->pc_gray.data(), n); }
  return (int) vo->pc_w.size();
}

}  // extern "C"
