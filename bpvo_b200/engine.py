"""Seam-level Python wrappers (Frame / PoseEstimator) over the C ABI -- what bpvo/vo.cc drives.

Method names mirror the reference classes: VisualOdometryFrame (bpvo/vo_frame.h) and
VisualOdometryPoseEstimator (bpvo/vo_pose_estimator.h).  numpy matrices are row-major 4x4 / 3x3;
they are transposed to the ABI's column-major float arrays at the boundary."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .types import AlgorithmParameters, Error, OptimizerStatistics, fill_cparams

FLAG_HOST_SOLVE = 1
FLAG_NO_GRAPHS = 2
FLAG_FAST_BLEND = 4
FLAG_TMA_DESCRIPTOR = 8


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _colmajor(M):
    return np.ascontiguousarray(np.asarray(M, dtype=np.float32).T).ravel()


def _from_colmajor(buf, n):
    return np.array(buf, dtype=np.float32).reshape(n, n).T.copy()


def _check(rc):
    if rc != 0:
        raise Error(f"[{rc}] {_capi.last_error()}")


class PinnedBuffer:
    """page-locked host array (cudaHostAlloc) for zero-staging uploads"""

    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self.ptr = _capi.lib().bpvo_b200_host_alloc(nbytes)
        if not self.ptr:
            raise Error("cudaHostAlloc failed: " + _capi.last_error())
        buf = (C.c_uint8 * nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            _capi.lib().bpvo_b200_host_free(self.ptr)
            self.ptr = None


class Context:
    """bpvo_b200_ctx: the pose estimator + device/stream binding (VisualOdometryPoseEstimator, vo.cc:98)."""

    def __init__(self, K, baseline, image_size, params: AlgorithmParameters, device_id: int = 0, flags: int = 0, _borrow=None):
        self._lib = _capi.lib()
        self.rows, self.cols = image_size
        self.params = params
        self.flags = int(flags)
        self._own = _borrow is None
        if _borrow is not None:
            self.h = C.c_void_p(_borrow)
            return
        p = AlgorithmParameters(**vars(params))
        p.numPyramidLevels = params.resolved_num_levels(self.rows, self.cols)
        self.params = p
        cp = fill_cparams(p, device_id, flags)
        self.h = C.c_void_p()
        _check(self._lib.bpvo_b200_create(C.byref(self.h), _fp(_colmajor(K)), float(baseline), self.rows, self.cols, C.byref(cp)))

    @property
    def channels(self):
        from .types import DescriptorType
        return {DescriptorType.kBitPlanes: 8, DescriptorType.kDescriptorFieldsFirstOrder: 5, DescriptorType.kIntensityAndGradient: 3}.get(self.params.descriptor, 1)

    def close(self):
        if getattr(self, "_own", False) and self.h:
            self._lib.bpvo_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def frame(self) -> "Frame":
        return Frame(self)

    # -- VisualOdometryPoseEstimator -------------------------------------------------------------
    def linearize(self, ref: "Frame", cur: "Frame", level: int, T, first_call_of_level: bool = True):
        H = np.zeros(36, np.float32)
        G = np.zeros(6, np.float32)
        f, s, nv = C.c_float(), C.c_float(), C.c_int32()
        _check(self._lib.bpvo_b200_linearize(self.h, ref.h, cur.h, level, _fp(_colmajor(T)), int(first_call_of_level),
                                             _fp(H), _fp(G), C.byref(f), C.byref(s), C.byref(nv)))
        return dict(f_norm=f.value, H=_from_colmajor(H, 6), G=G.copy(), sigma=s.value, n_valid=nv.value)

    def estimatePose(self, ref: "Frame", cur: "Frame", T_init):
        L = self.params.numPyramidLevels
        stats = (_capi.CStats * _capi.MAX_LEVELS)()
        T = np.zeros(16, np.float32)
        ev = C.c_int32()
        _check(self._lib.bpvo_b200_estimate_pose(self.h, ref.h, cur.h, _fp(_colmajor(T_init)), _fp(T), stats, C.byref(ev)))
        out = [OptimizerStatistics(s.numIterations, s.finalError, s.firstOrderOptimality, s.status) for s in stats[:L]]
        return _from_colmajor(T, 4), out, ev.value

    def debug_device_linearize(self, ref: "Frame", cur: "Frame", level: int, poses, grid_ctas: int = 0, cache_bytes: int = -1):
        """parity hook: len(poses) consecutive linearize() evaluations of `level` INSIDE the persistent kernel
        (bpvo_b200_debug_device_linearize); -> one dict per evaluation, same keys as linearize() + scale_path"""
        n = len(poses)
        buf = np.concatenate([_colmajor(T) for T in poses]).astype(np.float32)
        out = (_capi.CLinOut * n)()
        _check(self._lib.bpvo_b200_debug_device_linearize(self.h, ref.h, cur.h, level, _fp(buf), n, out, int(grid_ctas), int(cache_bytes)))
        return [dict(f_norm=o.f_norm, H=_from_colmajor(o.H, 6), G=np.array(o.G, np.float32), sigma=o.sigma, n_valid=o.n_valid,
                     n_good=o.n_good, scale_path=o.scale_path) for o in out]

    def set_trace(self, on: bool):
        _check(self._lib.bpvo_b200_debug_set_trace(self.h, int(on)))

    def get_trace(self, reset: bool = True):
        """rows {level, eval, f_norm, |dp|, max|G|, sigma, scale_path, status} of the GN loop since the last reset"""
        n = C.c_int32()
        _check(self._lib.bpvo_b200_debug_get_trace(self.h, None, 0, C.byref(n), 0))
        rows = np.zeros((max(n.value, 1), 8), np.float32)
        _check(self._lib.bpvo_b200_debug_get_trace(self.h, _fp(rows), n.value, C.byref(n), int(reset)))
        return rows[:n.value]

    def _vec(self, fn, dtype=np.float32):
        n = C.c_size_t(0)
        _check(fn(self.h, None, C.byref(n)))
        out = np.zeros(n.value, dtype)
        if n.value:
            cap = C.c_size_t(n.value)
            _check(fn(self.h, out.ctypes.data_as(fn.argtypes[1]), C.byref(cap)))
        return out

    def getWeights(self):
        return self._vec(self._lib.bpvo_b200_get_weights)

    def getResiduals(self):
        return self._vec(self._lib.bpvo_b200_get_residuals)

    def getValidFlags(self):
        return self._vec(self._lib.bpvo_b200_get_valid, np.uint8)

    def getFractionOfGoodPoints(self, thresh: float) -> float:
        f = C.c_float()
        _check(self._lib.bpvo_b200_fraction_good(self.h, float(thresh), C.byref(f)))
        return f.value

    # -- multi-GPU (point-sharded mode) ----------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        _check(_capi.lib().bpvo_b200_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, rank: int, nranks: int, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        _check(self._lib.bpvo_b200_comm_init(self.h, int(rank), int(nranks), buf))

    def comm_destroy(self):
        _check(self._lib.bpvo_b200_comm_destroy(self.h))

    def peer_export(self) -> bytes:
        """this rank's mailbox as a cudaIpcMemHandle_t (64 bytes) for the peer-memory mode"""
        buf = (C.c_uint8 * 64)()
        _check(self._lib.bpvo_b200_peer_export(self.h, buf))
        return bytes(buf)

    def peer_init(self, handles):
        """handles: the peer_export() results of ALL ranks in rank order (after comm_init)"""
        blob = b"".join(handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        _check(self._lib.bpvo_b200_peer_init(self.h, buf))

    def peer_set_min_points(self, n: int):
        _check(self._lib.bpvo_b200_peer_set_min_points(self.h, int(n)))

    def peer_init_distributed(self, dist):
        """convenience: exchange the handles with torch.distributed.all_gather_object and map the peers"""
        mine = self.peer_export()
        allh = [None] * dist.get_world_size()
        dist.all_gather_object(allh, mine)
        self.peer_init(allh)

    # -- measurement --------------------------------------------------------------------------------
    def set_solver_ctas(self, n: int):
        """throughput mode: SMs the on-device GN loop of this ctx occupies (0 = all)"""
        _check(self._lib.bpvo_b200_set_solver_ctas(self.h, int(n)))

    def set_profiling(self, on: bool):
        _check(self._lib.bpvo_b200_set_profiling(self.h, int(on)))

    def counters(self):
        c = _capi.CCounters()
        _check(self._lib.bpvo_b200_get_counters(self.h, C.byref(c)))
        return {k: getattr(c, k) for k, _ in _capi.CCounters._fields_}

    PHASES = ["P1_residuals", "sync1", "P2_select", "sync2", "P3_select", "sync3", "scale", "P4_reduce", "sync4",
              "final_sum", "solve", "other"]

    def phase_cycles(self, reset=True):
        buf = (C.c_longlong * 64)()
        _check(self._lib.bpvo_b200_get_phase_cycles(self.h, buf, int(reset)))
        out = dict(zip(self.PHASES, list(buf)[:len(self.PHASES)]))
        out["_bracket_hits"], out["_scale_estimates"] = buf[12], buf[13]
        out["_bracket_overflows"], out["_bracket_misses"] = buf[14], buf[15]
        out["_fine"] = list(buf)[16:64]
        return out

    def reset_counters(self):
        _check(self._lib.bpvo_b200_reset_counters(self.h))

    def synchronize(self):
        _check(self._lib.bpvo_b200_synchronize(self.h))

    def timer_start(self):
        _check(self._lib.bpvo_b200_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        _check(self._lib.bpvo_b200_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def last_level_evals(self):
        out = (C.c_int32 * _capi.MAX_LEVELS)()
        _check(self._lib.bpvo_b200_last_level_evals(self.h, out))
        return list(out[:self.params.numPyramidLevels])

    def last_level_us(self):
        out = (C.c_float * _capi.MAX_LEVELS)()
        _check(self._lib.bpvo_b200_last_level_us(self.h, out))
        return list(out[:self.params.numPyramidLevels])

    def level_phase_cycles(self, reset=True):
        """{level: {phase: cycles}} of the on-device GN loop while profiling was on"""
        buf = (C.c_longlong * (_capi.MAX_LEVELS * 16))()
        _check(self._lib.bpvo_b200_get_level_phase_cycles(self.h, buf, int(reset)))
        return {l: dict(zip(self.PHASES, list(buf)[l * 16:l * 16 + len(self.PHASES)])) for l in range(self.params.numPyramidLevels)}

    def time_linearize(self, ref, cur, level, T, iters=20, flush_l2=True) -> float:
        ms = C.c_float()
        _check(self._lib.bpvo_b200_time_linearize(self.h, ref.h, cur.h, level, _fp(_colmajor(T)), iters, int(flush_l2), C.byref(ms)))
        return ms.value


class Frame:
    """bpvo_b200_frame: VisualOdometryFrame (bpvo/vo_frame.h)."""

    def __init__(self, ctx: Context, _borrow=None):
        self.ctx = ctx
        self._lib = ctx._lib
        self._own = _borrow is None
        if _borrow is not None:
            self.h = C.c_void_p(_borrow)
        else:
            self.h = C.c_void_p()
            _check(self._lib.bpvo_b200_frame_create(ctx.h, C.byref(self.h)))

    def close(self):
        if getattr(self, "_own", False) and self.h and self.ctx.h:
            self._lib.bpvo_b200_frame_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setData(self, image, disparity):
        if image is None or disparity is None:
            _check(self._lib.bpvo_b200_frame_set_data(self.h, None, None))
        image = np.ascontiguousarray(image, dtype=np.uint8)
        disparity = np.ascontiguousarray(disparity, dtype=np.float32)
        assert image.shape == (self.ctx.rows, self.ctx.cols) and disparity.shape == image.shape
        _check(self._lib.bpvo_b200_frame_set_data(self.h, image.ctypes.data, disparity.ctypes.data))

    def setTemplate(self):
        _check(self._lib.bpvo_b200_frame_set_template(self.h))

    def hasTemplate(self):
        return bool(self._lib.bpvo_b200_frame_has_template(self.h))

    def empty(self):
        return bool(self._lib.bpvo_b200_frame_empty(self.h))

    def clear(self):
        _check(self._lib.bpvo_b200_frame_clear(self.h))

    def numLevels(self):
        return self._lib.bpvo_b200_frame_num_levels(self.h)

    def level_size(self, l):
        r, c = C.c_int32(), C.c_int32()
        _check(self._lib.bpvo_b200_frame_level_size(self.h, l, C.byref(r), C.byref(c)))
        return r.value, c.value

    def numPoints(self, l):
        n = C.c_int32()
        _check(self._lib.bpvo_b200_frame_num_points(self.h, l, C.byref(n)))
        return n.value

    def points(self, l):
        n = self.numPoints(l)
        out = np.zeros((n, 4), np.float32)
        if n:
            _check(self._lib.bpvo_b200_frame_get_points(self.h, l, _fp(out)))
        return out

    def pyramid(self, l):
        r, c = self.level_size(l)
        out = np.zeros((r, c), np.uint8)
        _check(self._lib.bpvo_b200_frame_get_pyramid(self.h, l, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def descriptor(self, l):
        r, c = self.level_size(l)
        ch = self.ctx.channels
        out = np.zeros((ch, r, c), np.float32)
        n = C.c_int32()
        _check(self._lib.bpvo_b200_frame_get_descriptor(self.h, l, _fp(out), C.byref(n)))
        return out

    def saliency(self, l):
        r, c = self.level_size(l)
        out = np.zeros((r, c), np.float32)
        _check(self._lib.bpvo_b200_frame_get_saliency(self.h, l, _fp(out)))
        return out

    def pixels(self, l):
        n, ch = self.numPoints(l), self.ctx.channels
        out = np.zeros((ch, n), np.float32)
        if n:
            _check(self._lib.bpvo_b200_frame_get_pixels(self.h, l, _fp(out)))
        return out

    def jacobians(self, l):
        n, ch = self.numPoints(l), self.ctx.channels
        out = np.zeros((ch, n, 6), np.float32)
        if n:
            _check(self._lib.bpvo_b200_frame_get_jacobians(self.h, l, _fp(out)))
        return out

    def point_inds(self, l):
        n = self.numPoints(l)
        out = np.zeros(n, np.int32)
        if n:
            _check(self._lib.bpvo_b200_frame_get_point_inds(self.h, l, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def normalization(self, l):
        out = np.zeros(16, np.float32)
        _check(self._lib.bpvo_b200_frame_get_normalization(self.h, l, _fp(out)))
        return _from_colmajor(out, 4)
