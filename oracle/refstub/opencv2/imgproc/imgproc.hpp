// Stand-ins for the cv::imgproc calls on the path.  pyrDown (u8) and GaussianBlur (5x5, CV_32F) are implemented in
// cv_stub_impl.cc with the formulas pinned bit-exact / to 1 ulp against cv2 4.13 golden vectors; everything else throws.
#pragma once
#include <opencv2/core/core.hpp>
namespace cv {
void pyrDown(const Mat& src, Mat& dst);
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0);
inline void cvtColor(const Mat&, Mat&, int) { throw std::logic_error("refstub: cv::cvtColor is not available"); }
inline void Laplacian(const Mat&, Mat&, int, int = 1) { throw std::logic_error("refstub: cv::Laplacian is not available"); }
}
