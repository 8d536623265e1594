/*
 * bpvo/vo_frame.h -- GPU-seam replacement of the reference header of the same name.
 *
 * Drop-in for halismai/bpvo: put this directory BEFORE the reference root on the include path, compile the reference's own
 * bpvo/vo.cc UNCHANGED together with integration/vo_b200_seam.cc, link libbpvo_b200.so -- VisualOdometry::addFrame() then runs
 * its key-frame state machine on the host and everything below it on the B200 (Option A).
 *
 * Same class name, same public methods with the same meaning as the reference's VisualOdometryFrame
 * (reference: bpvo/vo_frame.h:33-98, bpvo/vo_frame.cc:13-93); the private part holds a C-ABI handle instead of the
 * DenseDescriptorPyramid / TemplateData objects.  Only what bpvo/vo.cc calls is provided.
 */
#ifndef BPVO_VO_FRAME_H
#define BPVO_VO_FRAME_H

#include <bpvo/types.h>
#include <memory>
#include <vector>

namespace cv { class Mat; }
struct bpvo_b200_frame;

namespace bpvo {

namespace b200 { struct Group; }

/** what vo.cc reads through getTemplateDataAtLevel(): numPoints(), points(), warp().getImagePoint()
 *  (reference: bpvo/template_data.h:68-75, bpvo/rigid_body_warp.h:123-128) */
class TemplateData
{
 public:
  typedef typename EigenAlignedContainer<Point>::type PointVector;    // = VisualOdometry::PointVector (bpvo/vo.h)

  struct Warp {
    Matrix33 K;
    /** K * X.head<3>(), perspective division (rigid_body_warp.h:123-128) */
    inline ImagePoint getImagePoint(const Point& X) const
    {
      Eigen::Vector3f x = K * X.template head<3>();
      float z_i = 1.0f / x.z();
      return ImagePoint(z_i * x[0], z_i * x[1]);
    }
  };

  inline int numPoints() const { return static_cast<int>(points().size()); }
  /** the 3-D template points of this level, downloaded from the device on first use after a setTemplate() */
  const PointVector& points() const;
  inline const Warp& warp() const { return _warp; }

 private:
  friend class VisualOdometryFrame;
  const bpvo_b200_frame* _frame = nullptr;
  int _level = 0;
  Warp _warp;
  mutable bool _fresh = false;
  mutable PointVector _points;
}; // TemplateData

class VisualOdometryFrame
{
 public:
  VisualOdometryFrame(const Matrix33& K, float b, const AlgorithmParameters&);
  ~VisualOdometryFrame();

  VisualOdometryFrame(const VisualOdometryFrame&) = delete;
  VisualOdometryFrame& operator=(const VisualOdometryFrame&) = delete;

  /** copies both images to the device, builds the pyramid and the dense descriptors (vo_frame.cc:48-55) */
  void setData(const cv::Mat& image, const cv::Mat& disparity);
  /** TemplateData::setData for every level >= maxTestLevel, on the device (vo_frame.cc:61-93) */
  void setTemplate();

  inline void clear() { _has_data = false; _has_template = false; sync_flags(); }
  inline bool empty() { return !_has_data; }
  inline bool hasTemplate() const { return _has_template; }

  const TemplateData* getTemplateDataAtLevel(size_t) const;
  int numLevels() const;
  /** the raw input image (host copy, used by vo.cc to colour the point cloud) */
  const cv::Mat* imagePointer() const;

  /** the device-side frame (for VisualOdometryPoseEstimator) */
  inline const bpvo_b200_frame* handle() const { return _h; }

 private:
  void sync_flags();

  Matrix33 _K;
  float _b;
  AlgorithmParameters _params;
  bool _has_data;
  bool _has_template;
  std::shared_ptr<b200::Group> _group;      // the engine context shared with the pose estimator and the sibling frames
  bpvo_b200_frame* _h;
  UniquePointer<cv::Mat> _image;
  mutable std::vector<TemplateData> _tdata;
}; // VisualOdometryFrame

}; // bpvo

#endif // BPVO_VO_FRAME_H
