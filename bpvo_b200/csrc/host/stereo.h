// host/stereo.h -- C++ mirror of bpvo::StereoAlgorithm (utils/stereo_algorithm.h:14-37) for its default algorithm,
// "BlockMatching" (OpenCV's StereoBM as utils/stereo_algorithm.cc:67-111 configures and runs it), on the GPU through the
// bpvo_b200_stereo_* entry points of include/bpvo_b200.h.  Header only; same method names, errors as bpvo_b200::Error.
//
// The reference constructs it from a ConfigFile; here the caller fills StereoParameters (the same keys, the same defaults:
// bpvo_b200_stereo_default_params) from whatever configuration source it uses, e.g.
//     bpvo_b200::StereoParameters sp; bpvo_b200_stereo_default_params(&sp);
//     sp.numberOfDisparities = cf.get<int>("numberOfDisparities"); sp.SADWindowSize = cf.get<int>("SADWindowSize", 15); ...
#ifndef BPVO_B200_HOST_STEREO_H
#define BPVO_B200_HOST_STEREO_H

#include "vo.h"

namespace bpvo_b200 {

typedef bpvo_b200_stereo_params StereoParameters;

class StereoAlgorithm {
 public:
  StereoAlgorithm(ImageSize size, const StereoParameters& p) : _s(nullptr), _size(size) {
    if (bpvo_b200_stereo_create(&_s, size.rows, size.cols, &p) != BPVO_B200_OK) throw Error(bpvo_b200_last_error());
  }
  ~StereoAlgorithm() { bpvo_b200_stereo_destroy(_s); }
  StereoAlgorithm(const StereoAlgorithm&) = delete;
  StereoAlgorithm& operator=(const StereoAlgorithm&) = delete;

  // void run(const cv::Mat& left, const cv::Mat& right, cv::Mat& dmap): rows x cols u8 in, rows x cols f32 out
  // (host or device memory)
  void run(const uint8_t* left, const uint8_t* right, float* dmap) {
    if (bpvo_b200_stereo_run(_s, left, right, dmap, nullptr) != BPVO_B200_OK) throw Error(bpvo_b200_last_error());
  }
  float getInvalidValue() const { return bpvo_b200_stereo_invalid_value(_s); }      // the reference's short(minDisparity - 1) / 16.0f
  float filteredValue() const { return bpvo_b200_stereo_filtered_value(_s); }       // what invalid pixels hold in dmap: minDisparity - 1
  ImageSize imageSize() const { return _size; }
  bpvo_b200_stereo* handle() const { return _s; }

 private:
  bpvo_b200_stereo* _s;
  ImageSize _size;
};

}  // namespace bpvo_b200

#endif
