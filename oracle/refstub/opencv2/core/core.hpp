// Minimal stand-in for <opencv2/core/core.hpp>: a continuous, reference-counted cv::Mat with the members the
// reference's hot-path sources touch.  Test infrastructure only (see refstub/Eigen/Core).
#pragma once
// OpenCV 2.4's core/types_c.h includes <math.h>; with libstdc++ >= 6 that wrapper brings the float overloads of
// fabs() into the global namespace, which decides how imgproc.cc's unqualified `fabs(float)` tails are evaluated
// (float, as restated by the oracle).  Older toolchains resolve them to ::fabs(double) and differ by 1 ulp there.
#include <math.h>
#include <assert.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>

#define CV_8U 0
#define CV_32F 5
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) -1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_BGR2GRAY 6
#define CV_BGRA2GRAY 10

namespace cv {

struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
inline bool operator==(const Size& a, const Size& b) { return a.width == b.width && a.height == b.height; }

template <class T> struct DataType;
template <> struct DataType<uint8_t> { enum { type = CV_8UC1 }; };
template <> struct DataType<float> { enum { type = CV_32FC1 }; };

class Mat {
 public:
  int rows = 0, cols = 0;
  uint8_t* data = nullptr;
  Mat() {}
  Mat(Size s, int type) { create(s.height, s.width, type); }
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* ext) : rows(r), cols(c), data((uint8_t*) ext), type_(type) {}     // non-owning view
  void create(int r, int c, int type) {
    if (r == rows && c == cols && type == type_ && data) return;
    rows = r; cols = c; type_ = type;
    size_t bytes = (size_t) r * c * elemSize() + 64;
    void* p = nullptr; if (posix_memalign(&p, 64, bytes)) throw std::bad_alloc();
    own_.reset((uint8_t*) p, free); data = own_.get();
  }
  void create(Size s, int type) { create(s.height, s.width, type); }
  size_t elemSize() const { return (type_ & 7) == CV_32F ? 4 : 1; }
  int type() const { return type_; }
  int channels() const { return 1 + (type_ >> CV_CN_SHIFT); }
  bool isContinuous() const { return true; }
  bool empty() const { return data == nullptr; }
  Size size() const { return Size(cols, rows); }
  template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t) r * cols * elemSize()); }
  template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t) r * cols * elemSize()); }
  template <class T> T& at(int r, int c) { return ptr<T>(r)[c]; }
  template <class T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
  // OpenCV's copyTo takes an OutputArray (a const reference wrapper): a const destination is legal there
  void copyTo(const Mat& dst_) const { Mat& dst = const_cast<Mat&>(dst_); dst.create(rows, cols, type_); if (data) memcpy(dst.data, data, (size_t) rows * cols * elemSize()); }
  Mat clone() const { Mat m; copyTo(m); return m; }
  void convertTo(Mat& dst, int type) const;    // only CV_8U -> CV_32F and identity (cv_stub_impl.cc)
 protected:
  int type_ = CV_8UC1;
  std::shared_ptr<uint8_t> own_;
};

template <class T, size_t N = 1024> class AutoBuffer {      // histogram.h (only instantiated when DO_APPROX_MEDIAN is on)
 public:
  AutoBuffer() {}
  explicit AutoBuffer(size_t n) { allocate(n); }
  void allocate(size_t n) { buf_.reset(new T[n]); n_ = n; }
  size_t size() const { return n_; }
  operator T*() { return buf_.get(); }
  operator const T*() const { return buf_.get(); }
 private:
  std::unique_ptr<T[]> buf_; size_t n_ = 0;
};

template <class T> class Mat_ : public Mat {
 public:
  Mat_() { type_ = DataType<T>::type; }
  Mat_(int r, int c, T* ext) : Mat(r, c, DataType<T>::type, ext) {}
  Mat_(const Mat& m) : Mat(m) {}
  Mat_& operator=(const Mat& m) { Mat::operator=(m); return *this; }
  void create(int r, int c) { Mat::create(r, c, DataType<T>::type); }
  void convertTo(Mat& dst, int type) const { Mat::convertTo(dst, type); }
};

}  // namespace cv
