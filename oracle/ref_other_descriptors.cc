// Link-time stand-ins for the reference's OUT-OF-SCOPE descriptor classes (gradient / descriptor fields / LATCH /
// central difference / Laplacian).  DenseDescriptor::Create (bpvo/dense_descriptor.cc:38-90) names them, so their
// out-of-line members must exist for oracle/_ref to link; none of them is on the hot path, all of them throw.
#include <bpvo/central_difference_descriptor.h>
#include <bpvo/gradient_descriptor.h>
#include <bpvo/latch_descriptor.h>

#include <stdexcept>

namespace bpvo {
class LATCHDescriptorExtractorImpl {};
static void nope() { throw std::logic_error("oracle/_ref: this DescriptorType is outside the hot path and not built"); }

GradientDescriptor::GradientDescriptor(float s) : _rows(0), _cols(0), _sigma(s) {}
GradientDescriptor::GradientDescriptor(const GradientDescriptor& o) : DenseDescriptor(o), _rows(o._rows), _cols(o._cols), _sigma(o._sigma), _channels(o._channels) {}
GradientDescriptor::~GradientDescriptor() {}
void GradientDescriptor::compute(const cv::Mat&) { nope(); }
void LaplacianDescriptor::compute(const cv::Mat&) { nope(); }
DescriptorFields::DescriptorFields(float a, float b) : _rows(0), _cols(0), _sigma1(a), _sigma2(b) {}
DescriptorFields::DescriptorFields(const DescriptorFields& o) : DenseDescriptor(o), _rows(o._rows), _cols(o._cols), _sigma1(o._sigma1), _sigma2(o._sigma2), _channels(o._channels) {}
DescriptorFields::~DescriptorFields() {}
void DescriptorFields::compute(const cv::Mat&) { nope(); }
DescriptorFields2ndOrder::DescriptorFields2ndOrder(float a, float b) : _rows(0), _cols(0), _sigma1(a), _sigma2(b) {}
DescriptorFields2ndOrder::DescriptorFields2ndOrder(const DescriptorFields2ndOrder& o) : DenseDescriptor(o), _rows(o._rows), _cols(o._cols), _sigma1(o._sigma1), _sigma2(o._sigma2), _channels(o._channels) {}
DescriptorFields2ndOrder::~DescriptorFields2ndOrder() {}
void DescriptorFields2ndOrder::compute(const cv::Mat&) { nope(); }
LatchDescriptor::LatchDescriptor(int, bool, int) : _rows(0), _cols(0) {}
LatchDescriptor::LatchDescriptor(const LatchDescriptor& o) : DenseDescriptor(o), _rows(o._rows), _cols(o._cols), _channels(o._channels) {}
LatchDescriptor::~LatchDescriptor() {}
void LatchDescriptor::compute(const cv::Mat&) { nope(); }
CentralDifferenceDescriptor::CentralDifferenceDescriptor(int r, float a, float b) : _radius(r), _sigma_before(a), _sigma_after(b), _rows(0), _cols(0) {}
CentralDifferenceDescriptor::CentralDifferenceDescriptor(const CentralDifferenceDescriptor& o)
    : DenseDescriptor(o), _radius(o._radius), _sigma_before(o._sigma_before), _sigma_after(o._sigma_after), _rows(o._rows), _cols(o._cols), _channels(o._channels) {}
CentralDifferenceDescriptor::~CentralDifferenceDescriptor() {}
void CentralDifferenceDescriptor::compute(const cv::Mat&) { nope(); }
}  // namespace bpvo
