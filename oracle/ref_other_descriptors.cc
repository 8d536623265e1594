// Link-time stand-ins for the reference's OUT-OF-SCOPE descriptor classes (LATCH / central difference; the gradient,
// descriptor-fields and Laplacian classes are the reference's own bpvo/gradient_descriptor.cc since round 2).  DenseDescriptor::Create (bpvo/dense_descriptor.cc:38-90) names them, so their
// out-of-line members must exist for oracle/_ref to link; none of them is on the hot path, all of them throw.
#include <bpvo/central_difference_descriptor.h>
#include <bpvo/latch_descriptor.h>

#include <stdexcept>

namespace bpvo {
class LATCHDescriptorExtractorImpl {};
static void nope() { throw std::logic_error("oracle/_ref: this DescriptorType is outside the hot path and not built"); }

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
