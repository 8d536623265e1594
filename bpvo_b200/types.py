"""Python mirror of the reference's public value types (bpvo/types.h:125-566).

Names, numeric enum values and constructor defaults are the reference's (bpvo/types.cc:31-66) so
that code written against `bpvo::AlgorithmParameters` / `bpvo::Result` reads the same here.
"""
from __future__ import annotations

import ctypes
import enum
from dataclasses import dataclass, field, fields
from typing import List, Optional

import numpy as np


class LossFunctionType(enum.IntEnum):        # bpvo/types.h:125-130
    kHuber = 0x10
    kTukey = 0x11
    kL2 = 0x12


class VerbosityType(enum.IntEnum):           # bpvo/types.h:132-138
    kIteration = 0x20
    kFinal = 0x21
    kSilent = 0x22
    kDebug = 0x23


class DescriptorType(enum.IntEnum):          # bpvo/types.h:140-150 (only the two on the hot path are implemented)
    kIntensity = 0x30
    kIntensityAndGradient = 0x31
    kDescriptorFieldsFirstOrder = 0x32
    kDescriptorFieldsSecondOrder = 0x33
    kLatch = 0x34
    kCentralDifference = 0x35
    kLaplacian = 0x36
    kBitPlanes = 0x37


class GradientEstimationType(enum.IntEnum):  # bpvo/types.h:152-156
    kCentralDifference_3 = 0
    kCentralDifference_5 = 1


class InterpolationType(enum.IntEnum):       # bpvo/types.h:158-164
    kLinear = 0
    kCosine = 1
    kCubic = 2
    kCubicHermite = 3


class PoseEstimationStatus(enum.IntEnum):    # bpvo/types.h:399-406
    kParameterTolReached = 0x30
    kFunctionTolReached = 0x31
    kGradientTolReached = 0x32
    kMaxIterations = 0x33
    kSolverError = 0x34


class KeyFramingReason(enum.IntEnum):        # bpvo/types.h:411-418
    kLargeTranslation = 0x40
    kLargeRotation = 0x41
    kSmallFracOfGoodPoints = 0x42
    kNoKeyFraming = 0x43
    kFirstFrame = 0x44


@dataclass
class AlgorithmParameters:
    """bpvo::AlgorithmParameters, hot-path subset, ctor defaults of bpvo/types.cc:31-66."""
    numPyramidLevels: int = -1
    minImageDimensionForPyramid: int = 40
    sigmaPriorToCensusTransform: float = -1.0
    sigmaBitPlanes: float = 0.5
    maxIterations: int = 50
    parameterTolerance: float = 1e-7
    functionTolerance: float = 1e-6
    gradientTolerance: float = 1e-8
    relaxTolerancesForCoarseLevels: bool = True
    gradientEstimation: int = GradientEstimationType.kCentralDifference_3
    interp: int = InterpolationType.kLinear
    lossFunction: int = LossFunctionType.kTukey
    descriptor: int = DescriptorType.kIntensity
    verbosity: int = VerbosityType.kIteration
    minTranslationMagToKeyFrame: float = 0.15
    minRotationMagToKeyFrame: float = 5.0
    maxFractionOfGoodPointsToKeyFrame: float = 0.6
    goodPointThreshold: float = 0.85
    minNumPixelsForNonMaximaSuppression: int = 320 * 240
    nonMaxSuppRadius: int = 1
    minNumPixelsToWork: int = 256
    minSaliency: float = 0.1
    minValidDisparity: float = 0.001
    maxValidDisparity: float = 512.0
    maxTestLevel: int = 0
    withNormalization: bool = True
    dfSigma1: float = 0.75
    dfSigma2: float = 1.75

    def resolved_num_levels(self, rows: int, cols: int) -> int:
        """auto pyramid depth (bpvo/vo.cc:101-104)."""
        if self.numPyramidLevels > 0:
            return int(self.numPyramidLevels)
        import math
        # std::round (half away from zero), as the reference
        v = math.log2(min(rows, cols) / float(self.minImageDimensionForPyramid))
        return 1 + int(math.floor(abs(v) + 0.5) * (1 if v >= 0 else -1))


class CParams(ctypes.Structure):
    """`bpvo_b200_params` of include/bpvo_b200.h."""
    _fields_ = [
        ("numPyramidLevels", ctypes.c_int32), ("minImageDimensionForPyramid", ctypes.c_int32),
        ("sigmaPriorToCensusTransform", ctypes.c_float), ("sigmaBitPlanes", ctypes.c_float),
        ("maxIterations", ctypes.c_int32), ("parameterTolerance", ctypes.c_float),
        ("functionTolerance", ctypes.c_float), ("gradientTolerance", ctypes.c_float),
        ("relaxTolerancesForCoarseLevels", ctypes.c_int32), ("gradientEstimation", ctypes.c_int32),
        ("interp", ctypes.c_int32), ("lossFunction", ctypes.c_int32), ("descriptor", ctypes.c_int32),
        ("verbosity", ctypes.c_int32), ("minTranslationMagToKeyFrame", ctypes.c_float),
        ("minRotationMagToKeyFrame", ctypes.c_float), ("maxFractionOfGoodPointsToKeyFrame", ctypes.c_float),
        ("goodPointThreshold", ctypes.c_float), ("minNumPixelsForNonMaximaSuppression", ctypes.c_int32),
        ("nonMaxSuppRadius", ctypes.c_int32), ("minNumPixelsToWork", ctypes.c_int32),
        ("minSaliency", ctypes.c_float), ("minValidDisparity", ctypes.c_float),
        ("maxValidDisparity", ctypes.c_float), ("maxTestLevel", ctypes.c_int32),
        ("withNormalization", ctypes.c_int32),
        # engine options: x0 = device_id, x1 = flags (BPVO_B200_FLAG_*)
        ("x0", ctypes.c_int32), ("x1", ctypes.c_int32),
        ("dfSigma1", ctypes.c_float), ("dfSigma2", ctypes.c_float),
    ]


def fill_cparams(p: AlgorithmParameters, x0: int = 0, x1: int = 0) -> CParams:
    c = CParams()
    for f in fields(p):
        setattr(c, f.name, type(getattr(c, f.name))(getattr(p, f.name)))
    c.x0, c.x1 = int(x0), int(x1)
    return c


@dataclass
class OptimizerStatistics:                    # bpvo/types.h:444-482, defaults types.cc:306-310
    numIterations: int = 0
    finalError: float = -1.0
    firstOrderOptimality: float = -1.0
    status: int = PoseEstimationStatus.kSolverError


@dataclass
class PointCloud:                             # bpvo/point_cloud.h (xyzw, weight, gray colour)
    points: np.ndarray
    weights: np.ndarray
    gray: np.ndarray
    pose: np.ndarray


@dataclass
class Result:                                 # bpvo/types.h:496-566
    pose: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    covariance: np.ndarray = field(default_factory=lambda: np.eye(6, dtype=np.float32))   # never computed by the reference (Q11)
    optimizerStatistics: List[OptimizerStatistics] = field(default_factory=list)
    isKeyFrame: bool = False
    keyFramingReason: int = KeyFramingReason.kNoKeyFraming
    pointCloud: Optional[PointCloud] = None
    numFunEvals: int = 0                      # extra: linearize() calls inside this addFrame (GN iterations)


class Error(RuntimeError):
    """bpvo::Error (bpvo/utils.h:211-220)."""
