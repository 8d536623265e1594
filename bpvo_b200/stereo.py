"""Host-side mirror of bpvo::StereoAlgorithm (utils/stereo_algorithm.h:14-37) for its default algorithm, "BlockMatching":
OpenCV's StereoBM as utils/stereo_algorithm.cc:67-111 configures and runs it, on the GPU (bpvo_b200/csrc/stereo.cu, reached
through the bpvo_b200_stereo_* entry points of include/bpvo_b200.h).  No compute here and no CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .engine import _check
from .types import Error

# the keys StereoAlgorithm::Impl reads from the ConfigFile for "BlockMatching" (stereo_algorithm.cc:70-84) and their defaults
_BM_KEYS = ("preFilterType", "preFilterSize", "preFilterCap", "SADWindowSize", "minDisparity", "numberOfDisparities",
            "textureThreshold", "uniquenessRatio", "speckleWindowSize", "speckleRange", "trySmallerWindows", "disp12MaxDiff")


class StereoAlgorithm:
    """StereoAlgorithm(const ConfigFile&): `config` is a mapping with the reference's keys (e.g. the parsed conf/kitti.cfg);
    keyword arguments override it.  `numberOfDisparities` must be provided, as in the reference (stereo_algorithm.cc:75)."""

    def __init__(self, image_size, config=None, device_id: int = 0, **kw):
        cfg = dict(config or {})
        cfg.update(kw)
        alg = str(cfg.get("StereoAlgorithm", "BlockMatching"))
        if alg.lower() not in ("blockmatching", "bm"):                       # icompare, stereo_algorithm.cc:65
            if alg.lower() in ("sgbm", "semiglobalblockmatching", "sgm", "semiglobalmatching", "rsgm"):
                raise Error(f"StereoAlgorithm {alg}: only BlockMatching is on the accelerated path")
            raise Error(f"Unknown stereo algorithm {alg}\n")                 # stereo_algorithm.cc:87
        if "numberOfDisparities" not in cfg:
            raise Error("no key numberOfDisparities")                         # ConfigFile::get without a default throws
        self._lib = _capi.lib()
        p = _capi.CStereoParams()
        self._lib.bpvo_b200_stereo_default_params(C.byref(p))
        for k in _BM_KEYS:
            if k in cfg:
                setattr(p, k, int(cfg[k]))
        p.device_id = device_id
        self.params = p
        self.rows, self.cols = int(image_size[0]), int(image_size[1])
        self.h = C.c_void_p()
        _check(self._lib.bpvo_b200_stereo_create(C.byref(self.h), self.rows, self.cols, C.byref(p)))

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self._lib.bpvo_b200_stereo_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    def run(self, left, right, want_fixed_point: bool = False):
        """void run(const cv::Mat& left, const cv::Mat& right, cv::Mat& dmap) -> dmap (float32, pixels); with
        want_fixed_point also OpenCV's CV_16S map (what cvFindStereoCorrespondenceBM itself returns)"""
        left = np.ascontiguousarray(left, np.uint8); right = np.ascontiguousarray(right, np.uint8)
        if left.shape != (self.rows, self.cols) or right.shape != left.shape:
            raise ValueError("image size mismatch")
        dmap = np.empty(left.shape, np.float32)
        d16 = np.empty(left.shape, np.int16) if want_fixed_point else None
        _check(self._lib.bpvo_b200_stereo_run(self.h, left.ctypes.data, right.ctypes.data, dmap.ctypes.data, d16.ctypes.data if want_fixed_point else None))
        return (dmap, d16) if want_fixed_point else dmap

    def run_raw(self, left_ptr: int, right_ptr: int, dmap_ptr: int, disp16_ptr: int = 0):
        """same, from raw host or device addresses: no numpy work in a timed path"""
        _check(self._lib.bpvo_b200_stereo_run(self.h, left_ptr, right_ptr, dmap_ptr or None, disp16_ptr or None))

    def getInvalidValue(self) -> float:
        """float getInvalidValue() const, with the reference's arithmetic (stereo_algorithm.cc:138-146): short(minDisparity - 1) / 16,
        i.e. -0.0625 by default -- NOT what invalid pixels hold in the map; that is filteredValue()"""
        return float(self._lib.bpvo_b200_stereo_invalid_value(self.h))

    def filteredValue(self) -> float:
        """the value invalid pixels carry in the disparity map: minDisparity - 1"""
        return float(self._lib.bpvo_b200_stereo_filtered_value(self.h))

    def prefiltered(self):
        a = np.empty((self.rows, self.cols), np.uint8); b = np.empty_like(a)
        _check(self._lib.bpvo_b200_stereo_get_prefiltered(self.h, a.ctypes.data, b.ctypes.data))
        return a, b

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        _check(self._lib.bpvo_b200_stereo_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def launches(self) -> int:
        return int(self._lib.bpvo_b200_stereo_launches(self.h))
