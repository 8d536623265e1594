"""bpvo::VisualOdometry (bpvo/vo.h:33-107) for Python, bound to the VisualOdometry-level C ABI.

    vo = VisualOdometry(K, baseline, (rows, cols), AlgorithmParameters(...))
    result = vo.addFrame(image_u8, disparity_f32)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .engine import Context, Frame, _check, _colmajor, _fp, _from_colmajor
from .types import AlgorithmParameters, OptimizerStatistics, PointCloud, Result, fill_cparams


class VisualOdometry:
    def __init__(self, K, baseline: float, image_size, params: AlgorithmParameters = None, device_id: int = 0, flags: int = 0):
        self._lib = _capi.lib()
        self.params = params or AlgorithmParameters()
        self.rows, self.cols = (image_size.rows, image_size.cols) if hasattr(image_size, "rows") else image_size
        cp = fill_cparams(self.params, device_id, flags)
        self.h = C.c_void_p()
        _check(self._lib.bpvo_b200_vo_create(C.byref(self.h), _fp(_colmajor(K)), float(baseline), self.rows, self.cols, C.byref(cp)))
        self._res = _capi.CResult()
        p = AlgorithmParameters(**vars(self.params))
        p.numPyramidLevels = self.params.resolved_num_levels(self.rows, self.cols)
        self._ctx = Context(K, baseline, (self.rows, self.cols), p, _borrow=self._lib.bpvo_b200_vo_ctx(self.h))

    def close(self):
        if getattr(self, "h", None):
            self._lib.bpvo_b200_vo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def ctx(self) -> Context:
        return self._ctx

    def ref_frame(self) -> Frame:
        return Frame(self._ctx, _borrow=self._lib.bpvo_b200_vo_ref_frame(self.h))

    def addFrame(self, image, disparity) -> Result:
        """Result addFrame(const uint8_t* image, const float* disparity) (bpvo/vo.h:75)."""
        if image is None or disparity is None:
            _check(self._lib.bpvo_b200_vo_add_frame(self.h, None, None, C.byref(self._res)))
        image = np.ascontiguousarray(image, dtype=np.uint8)
        disparity = np.ascontiguousarray(disparity, dtype=np.float32)
        if image.shape != (self.rows, self.cols) or disparity.shape != image.shape:
            raise ValueError("image / disparity size mismatch")
        return self.addFrameRaw(image.ctypes.data, disparity.ctypes.data)

    def addFrameRaw(self, image_ptr: int, disparity_ptr: int, want_cloud: bool = True, _stereo=None) -> Result:
        """same, from raw host addresses (e.g. PinnedBuffer.ptr): no numpy work in the timed path"""
        r = self._res
        if _stereo is not None:
            _check(self._lib.bpvo_b200_vo_add_stereo_frame(self.h, _stereo[0].h, image_ptr, _stereo[1], C.byref(r)))
        else:
            _check(self._lib.bpvo_b200_vo_add_frame(self.h, image_ptr, disparity_ptr, C.byref(r)))
        out = Result()
        out.pose = _from_colmajor(r.pose, 4)
        out.isKeyFrame = bool(r.isKeyFrame)
        out.keyFramingReason = r.keyFramingReason
        out.optimizerStatistics = [OptimizerStatistics(s.numIterations, s.finalError, s.firstOrderOptimality, s.status)
                                   for s in r.optimizerStatistics[:r.numLevels]]
        out.numFunEvals = r.numFunEvals
        if want_cloud and r.numPointCloud > 0:
            n = r.numPointCloud
            xyzw = np.zeros((n, 4), np.float32)
            w = np.zeros(n, np.float32)
            g = np.zeros(n, np.uint8)
            cnt = C.c_int32()
            _check(self._lib.bpvo_b200_vo_point_cloud(self.h, _fp(xyzw), _fp(w), g.ctypes.data_as(C.POINTER(C.c_uint8)), n, C.byref(cnt)))
            out.pointCloud = PointCloud(xyzw, w, g, self.trajectory()[-1])
        return out

    def addStereoFrame(self, stereo, left, right, want_cloud: bool = True) -> Result:
        """image pair in, pose out: `stereo` (bpvo_b200.stereo.StereoAlgorithm) produces the disparity map on the device and
        addFrame consumes it there (what utils/dataset.cc:133 + VisualOdometry::addFrame do on the host)"""
        left = np.ascontiguousarray(left, dtype=np.uint8); right = np.ascontiguousarray(right, dtype=np.uint8)
        if left.shape != (self.rows, self.cols) or right.shape != left.shape:
            raise ValueError("image size mismatch")
        return self.addFrameRaw(left.ctypes.data, 0, want_cloud, _stereo=(stereo, right.ctypes.data))

    def numPointsAtLevel(self, level: int = -1) -> int:
        n = C.c_int32()
        _check(self._lib.bpvo_b200_vo_num_points_at_level(self.h, level, C.byref(n)))
        return n.value

    def pointsAtLevel(self, level: int = -1) -> np.ndarray:
        n = self.numPointsAtLevel(level)
        out = np.zeros((n, 4), np.float32)
        if n:
            _check(self._lib.bpvo_b200_vo_points_at_level(self.h, level, _fp(out), n))
        return out

    def trajectory(self) -> np.ndarray:
        n = C.c_int32()
        _check(self._lib.bpvo_b200_vo_trajectory(self.h, None, 0, C.byref(n)))
        buf = np.zeros((max(n.value, 1), 16), np.float32)
        _check(self._lib.bpvo_b200_vo_trajectory(self.h, _fp(buf), n.value, C.byref(n)))
        return np.stack([b.reshape(4, 4).T for b in buf[:n.value]]) if n.value else np.zeros((0, 4, 4), np.float32)
