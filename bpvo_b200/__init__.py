"""bpvo_b200 -- B200-native dense-alignment engine behind bpvo's VisualOdometry::addFrame().

The compute path is hand-written sm_100a CUDA in libbpvo_b200.so (bpvo_b200/csrc), reached through the
C ABI of include/bpvo_b200.h.  This package is the Python host-side mirror of the reference's public
interface; it contains no compute and no CPU fallback."""
from .types import (AlgorithmParameters, DescriptorType, Error, GradientEstimationType, InterpolationType,
                    KeyFramingReason, LossFunctionType, OptimizerStatistics, PointCloud, PoseEstimationStatus,
                    Result, VerbosityType)

__all__ = ["AlgorithmParameters", "DescriptorType", "Error", "GradientEstimationType", "InterpolationType",
           "KeyFramingReason", "LossFunctionType", "OptimizerStatistics", "PointCloud", "PoseEstimationStatus",
           "Result", "VerbosityType", "VisualOdometry", "Context", "Frame", "StereoAlgorithm"]


def __getattr__(name):
    # the classes below load the CUDA library on first use
    if name == "VisualOdometry":
        from .vo import VisualOdometry
        return VisualOdometry
    if name == "StereoAlgorithm":
        from .stereo import StereoAlgorithm
        return StereoAlgorithm
    if name in ("Context", "Frame", "PinnedBuffer"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
