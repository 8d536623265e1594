"""ctypes binding of libbpvo_b200.so (include/bpvo_b200.h).  There is no fallback: if the CUDA
library is missing this raises, and if no CUDA device is present every create() fails."""
from __future__ import annotations

import ctypes as C
import os

from .types import CParams

_HERE = os.path.dirname(os.path.abspath(__file__))
# BPVO_B200_LIB: profiling builds of the SAME sources (e.g. libbpvo_b200_fine.so); the default is the product library
LIB_PATH = os.environ.get("BPVO_B200_LIB") or os.path.join(_HERE, "libbpvo_b200.so")
_LIB = None

MAX_LEVELS = 16


class CStats(C.Structure):
    _fields_ = [("numIterations", C.c_int32), ("finalError", C.c_float),
                ("firstOrderOptimality", C.c_float), ("status", C.c_int32)]


class CResult(C.Structure):
    _fields_ = [("pose", C.c_float * 16), ("isKeyFrame", C.c_int32), ("keyFramingReason", C.c_int32),
                ("numLevels", C.c_int32), ("optimizerStatistics", CStats * MAX_LEVELS),
                ("numFunEvals", C.c_int32), ("numPointCloud", C.c_int32)]


class CLinOut(C.Structure):
    _fields_ = [("H", C.c_float * 36), ("G", C.c_float * 6), ("f_norm", C.c_float), ("sigma", C.c_float),
                ("n_valid", C.c_int32), ("n_good", C.c_int32), ("solve_ok", C.c_int32), ("scale_path", C.c_int32)]


class CCounters(C.Structure):
    _fields_ = [("ms_upload", C.c_double), ("ms_pyramid", C.c_double), ("ms_descriptor", C.c_double),
                ("ms_template", C.c_double), ("ms_linearize", C.c_double), ("ms_total", C.c_double),
                ("launches", C.c_int64), ("linearize_calls", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("solve_calls", C.c_int64)]


class CStereoParams(C.Structure):          # bpvo_b200_stereo_params: the CvStereoBMState fields utils/stereo_algorithm.cc:67-85 sets
    _fields_ = [(n, C.c_int32) for n in ("numberOfDisparities", "SADWindowSize", "minDisparity", "preFilterType", "preFilterSize",
                                         "preFilterCap", "textureThreshold", "uniquenessRatio", "speckleWindowSize", "speckleRange",
                                         "trySmallerWindows", "disp12MaxDiff", "device_id")]


# name -> (restype, argtypes); exactly the symbols include/bpvo_b200.h declares
def _signatures():
    vp, fp, u8p, ip = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
    pp = C.POINTER(CParams)
    szp = C.POINTER(C.c_size_t)
    return {
        "bpvo_b200_version": (C.c_int, []),
        "bpvo_b200_last_error": (C.c_char_p, []),
        "bpvo_b200_default_params": (None, [pp]),
        "bpvo_b200_device_count": (C.c_int, []),
        "bpvo_b200_vo_create": (C.c_int, [C.POINTER(vp), fp, C.c_float, C.c_int, C.c_int, pp]),
        "bpvo_b200_vo_destroy": (C.c_int, [vp]),
        "bpvo_b200_vo_add_frame": (C.c_int, [vp, C.c_void_p, C.c_void_p, C.POINTER(CResult)]),
        "bpvo_b200_vo_num_points_at_level": (C.c_int, [vp, C.c_int, ip]),
        "bpvo_b200_vo_points_at_level": (C.c_int, [vp, C.c_int, fp, C.c_int]),
        "bpvo_b200_vo_trajectory": (C.c_int, [vp, fp, C.c_int, ip]),
        "bpvo_b200_vo_point_cloud": (C.c_int, [vp, fp, fp, u8p, C.c_int, ip]),
        "bpvo_b200_vo_ctx": (vp, [vp]),
        "bpvo_b200_vo_ref_frame": (vp, [vp]),
        "bpvo_b200_create": (C.c_int, [C.POINTER(vp), fp, C.c_float, C.c_int, C.c_int, pp]),
        "bpvo_b200_destroy": (C.c_int, [vp]),
        "bpvo_b200_frame_create": (C.c_int, [vp, C.POINTER(vp)]),
        "bpvo_b200_frame_destroy": (C.c_int, [vp]),
        "bpvo_b200_frame_set_data": (C.c_int, [vp, C.c_void_p, C.c_void_p]),
        "bpvo_b200_frame_set_template": (C.c_int, [vp]),
        "bpvo_b200_frame_has_template": (C.c_int, [vp]),
        "bpvo_b200_frame_empty": (C.c_int, [vp]),
        "bpvo_b200_frame_clear": (C.c_int, [vp]),
        "bpvo_b200_frame_num_levels": (C.c_int, [vp]),
        "bpvo_b200_frame_level_size": (C.c_int, [vp, C.c_int, ip, ip]),
        "bpvo_b200_frame_num_points": (C.c_int, [vp, C.c_int, ip]),
        "bpvo_b200_frame_get_points": (C.c_int, [vp, C.c_int, fp]),
        "bpvo_b200_frame_get_pyramid": (C.c_int, [vp, C.c_int, u8p]),
        "bpvo_b200_frame_get_descriptor": (C.c_int, [vp, C.c_int, fp, ip]),
        "bpvo_b200_frame_get_saliency": (C.c_int, [vp, C.c_int, fp]),
        "bpvo_b200_frame_get_pixels": (C.c_int, [vp, C.c_int, fp]),
        "bpvo_b200_frame_get_jacobians": (C.c_int, [vp, C.c_int, fp]),
        "bpvo_b200_frame_get_point_inds": (C.c_int, [vp, C.c_int, ip]),
        "bpvo_b200_frame_get_normalization": (C.c_int, [vp, C.c_int, fp]),
        "bpvo_b200_linearize": (C.c_int, [vp, vp, vp, C.c_int, fp, C.c_int, fp, fp, fp, fp, ip]),
        "bpvo_b200_estimate_pose": (C.c_int, [vp, vp, vp, fp, fp, C.POINTER(CStats), ip]),
        "bpvo_b200_get_weights": (C.c_int, [vp, fp, szp]),
        "bpvo_b200_get_residuals": (C.c_int, [vp, fp, szp]),
        "bpvo_b200_get_valid": (C.c_int, [vp, u8p, szp]),
        "bpvo_b200_debug_cache_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32)]),
        "bpvo_b200_debug_device_linearize": (C.c_int, [vp, vp, vp, C.c_int, fp, C.c_int, C.POINTER(CLinOut), C.c_int, C.c_int]),
        "bpvo_b200_debug_set_trace": (C.c_int, [vp, C.c_int]),
        "bpvo_b200_debug_get_trace": (C.c_int, [vp, fp, C.c_int, ip, C.c_int]),
        "bpvo_b200_point_cloud": (C.c_int, [vp, vp, C.c_void_p, ip]),
        "bpvo_b200_fraction_good": (C.c_int, [vp, C.c_float, fp]),
        "bpvo_b200_comm_unique_id": (C.c_int, [u8p]),
        "bpvo_b200_comm_init": (C.c_int, [vp, C.c_int, C.c_int, u8p]),
        "bpvo_b200_comm_destroy": (C.c_int, [vp]),
        "bpvo_b200_peer_export": (C.c_int, [vp, u8p]),
        "bpvo_b200_peer_init": (C.c_int, [vp, u8p]),
        "bpvo_b200_peer_set_min_points": (C.c_int, [vp, C.c_int]),
        "bpvo_b200_set_profiling": (C.c_int, [vp, C.c_int]),
        "bpvo_b200_set_solver_ctas": (C.c_int, [vp, C.c_int]),
        "bpvo_b200_get_counters": (C.c_int, [vp, C.POINTER(CCounters)]),
        "bpvo_b200_reset_counters": (C.c_int, [vp]),
        "bpvo_b200_get_phase_cycles": (C.c_int, [vp, C.POINTER(C.c_longlong), C.c_int]),
        "bpvo_b200_synchronize": (C.c_int, [vp]),
        "bpvo_b200_timer_start": (C.c_int, [vp]),
        "bpvo_b200_timer_stop": (C.c_int, [vp, fp]),
        "bpvo_b200_last_level_evals": (C.c_int, [vp, ip]),
        "bpvo_b200_last_level_us": (C.c_int, [vp, fp]),
        "bpvo_b200_get_level_phase_cycles": (C.c_int, [vp, C.POINTER(C.c_longlong), C.c_int]),
        "bpvo_b200_host_alloc": (C.c_void_p, [C.c_size_t]),
        "bpvo_b200_host_free": (None, [C.c_void_p]),
        "bpvo_b200_time_linearize": (C.c_int, [vp, vp, vp, C.c_int, fp, C.c_int, C.c_int, fp]),
        "bpvo_b200_stereo_default_params": (None, [C.POINTER(CStereoParams)]),
        "bpvo_b200_stereo_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(CStereoParams)]),
        "bpvo_b200_stereo_destroy": (C.c_int, [vp]),
        "bpvo_b200_stereo_run": (C.c_int, [vp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
        "bpvo_b200_stereo_invalid_value": (C.c_float, [vp]),
        "bpvo_b200_stereo_filtered_value": (C.c_float, [vp]),
        "bpvo_b200_stereo_get_prefiltered": (C.c_int, [vp, C.c_void_p, C.c_void_p]),
        "bpvo_b200_stereo_last_kernel_ms": (C.c_int, [vp, fp]),
        "bpvo_b200_stereo_launches": (C.c_longlong, [vp]),
        "bpvo_b200_vo_add_stereo_frame": (C.c_int, [vp, vp, C.c_void_p, C.c_void_p, C.POINTER(CResult)]),
    }


SIGNATURES = _signatures()


def lib():
    """Loads the CUDA library; raises if it has not been built (no CPU fallback exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m bpvo_b200.build` "
                          "(bpvo_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        f = getattr(L, name)          # AttributeError if the .so does not export a declared symbol
        f.restype, f.argtypes = res, args
    _LIB = L
    return L


def last_error() -> str:
    return lib().bpvo_b200_last_error().decode()
