#pragma once
#include <opencv2/core/core.hpp>
namespace cv {
// the reference only reaches these with sigma > 0 (census pre-blur, imsmooth): not available in the stub
inline void GaussianBlur(const Mat&, Mat&, Size, double, double = 0) { throw std::logic_error("refstub: cv::GaussianBlur is not available"); }
}
