"""ctypes binding of oracle/liboracle.so -- the CPU restatement of the reference.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under bpvo_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OrcParams(C.Structure):
    _fields_ = [
        ("numPyramidLevels", C.c_int32), ("minImageDimensionForPyramid", C.c_int32),
        ("sigmaPriorToCensusTransform", C.c_float), ("sigmaBitPlanes", C.c_float),
        ("maxIterations", C.c_int32), ("parameterTolerance", C.c_float),
        ("functionTolerance", C.c_float), ("gradientTolerance", C.c_float),
        ("relaxTolerancesForCoarseLevels", C.c_int32), ("gradientEstimation", C.c_int32),
        ("interp", C.c_int32), ("lossFunction", C.c_int32), ("descriptor", C.c_int32),
        ("verbosity", C.c_int32), ("minTranslationMagToKeyFrame", C.c_float),
        ("minRotationMagToKeyFrame", C.c_float), ("maxFractionOfGoodPointsToKeyFrame", C.c_float),
        ("goodPointThreshold", C.c_float), ("minNumPixelsForNonMaximaSuppression", C.c_int32),
        ("nonMaxSuppRadius", C.c_int32), ("minNumPixelsToWork", C.c_int32),
        ("minSaliency", C.c_float), ("minValidDisparity", C.c_float),
        ("maxValidDisparity", C.c_float), ("maxTestLevel", C.c_int32),
        ("withNormalization", C.c_int32), ("use_rcp", C.c_int32), ("num_threads", C.c_int32),
        ("dfSigma1", C.c_float), ("dfSigma2", C.c_float),
    ]


class OrcStats(C.Structure):
    _fields_ = [("numIterations", C.c_int32), ("finalError", C.c_float),
                ("firstOrderOptimality", C.c_float), ("status", C.c_int32)]


class OrcResult(C.Structure):
    _fields_ = [("pose", C.c_float * 16), ("isKeyFrame", C.c_int32), ("keyFramingReason", C.c_int32),
                ("numLevels", C.c_int32), ("stats", OrcStats * 16), ("numFunEvals", C.c_int32),
                ("numPointCloud", C.c_int32)]


class OrcStereoParams(C.Structure):
    _fields_ = [("numberOfDisparities", C.c_int), ("SADWindowSize", C.c_int), ("minDisparity", C.c_int),
                ("preFilterCap", C.c_int), ("textureThreshold", C.c_int), ("uniquenessRatio", C.c_int)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("bpvo_oracle.cc", "stereo_oracle.cc", "bpvo_oracle.h")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-s", "-B", "liboracle.so"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    vp, fp, u8p, ip = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
    sig = {
        "orc_default_params": (None, [C.POINTER(OrcParams)]),
        "orc_pyr_down": (None, [u8p, C.c_int, C.c_int, u8p]),
        "orc_gaussian_blur5": (None, [fp, C.c_int, C.c_int, C.c_float, fp]),
        "orc_gaussian_blur_f32": (None, [fp, C.c_int, C.c_int, C.c_int, C.c_double, fp]),
        "orc_census": (None, [u8p, C.c_int, C.c_int, u8p]),
        "orc_gaussian_blur3_u8": (None, [u8p, C.c_int, C.c_int, C.c_float, u8p]),
        "orc_descriptor": (C.c_int, [C.POINTER(OrcParams), u8p, C.c_int, C.c_int, fp]),
        "orc_saliency": (None, [fp, C.c_int, C.c_int, C.c_int, fp]),
        "orc_median": (C.c_float, [fp, C.c_size_t]),
        "orc_solve6": (C.c_int, [fp, fp, fp]),
        "orc_params_to_pose": (None, [fp, fp, fp]),
        "orc_frame_create": (vp, [fp, C.c_float, C.c_int, C.c_int, C.POINTER(OrcParams)]),
        "orc_frame_destroy": (None, [vp]),
        "orc_frame_set_data": (None, [vp, u8p, fp]),
        "orc_frame_set_template": (C.c_int, [vp]),
        "orc_frame_num_levels": (C.c_int, [vp]),
        "orc_frame_level_size": (None, [vp, C.c_int, ip, ip]),
        "orc_frame_pyramid": (u8p, [vp, C.c_int]),
        "orc_frame_descriptor": (fp, [vp, C.c_int, ip]),
        "orc_frame_saliency": (fp, [vp, C.c_int]),
        "orc_frame_num_points": (C.c_int, [vp, C.c_int]),
        "orc_frame_points": (fp, [vp, C.c_int]),
        "orc_frame_pixels": (fp, [vp, C.c_int]),
        "orc_frame_jacobians": (fp, [vp, C.c_int]),
        "orc_frame_point_inds": (ip, [vp, C.c_int]),
        "orc_frame_normalization": (None, [vp, C.c_int, fp]),
        "orc_estimator_create": (vp, [C.POINTER(OrcParams)]),
        "orc_estimator_destroy": (None, [vp]),
        "orc_linearize": (C.c_float, [vp, vp, vp, C.c_int, fp, C.c_int, fp, fp, fp]),
        "orc_estimator_num_residuals": (C.c_size_t, [vp]),
        "orc_estimator_residuals": (fp, [vp]),
        "orc_estimator_weights": (fp, [vp]),
        "orc_estimator_valid": (C.POINTER(C.c_uint16), [vp]),
        "orc_estimate_pose": (C.c_int, [vp, vp, vp, fp, fp, C.POINTER(OrcStats)]),
        "orc_fraction_good": (C.c_float, [vp, C.c_float]),
        "orc_estimator_set_trace": (None, [vp, C.c_int]),
        "orc_estimator_get_trace": (C.c_int, [vp, fp, C.c_int]),
        "orc_vo_create": (vp, [fp, C.c_float, C.c_int, C.c_int, C.POINTER(OrcParams)]),
        "orc_vo_destroy": (None, [vp]),
        "orc_vo_add_frame": (C.c_int, [vp, u8p, fp, C.POINTER(OrcResult)]),
        "orc_vo_num_points_at_level": (C.c_int, [vp, C.c_int]),
        "orc_vo_ref_frame": (vp, [vp]),
        "orc_vo_estimator": (vp, [vp]),
        "orc_vo_trajectory": (C.c_int, [vp, fp, C.c_int]),
        "orc_vo_point_cloud": (C.c_int, [vp, fp, fp, u8p, C.c_int]),
        "orc_last_error": (C.c_char_p, []),
        "orc_stereo_prefilter_xsobel": (None, [u8p, C.c_int, C.c_int, C.c_int, u8p]),
        "orc_stereo_bm": (C.c_int, [u8p, u8p, C.c_int, C.c_int, C.POINTER(OrcStereoParams), C.POINTER(C.c_int16), fp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _LIB = L
    return L


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _colmajor(M):
    """row-major numpy matrix -> flat column-major float32 (what crosses the C API)."""
    return np.ascontiguousarray(np.asarray(M, dtype=np.float32).T).ravel()


def _from_colmajor(buf, n):
    return np.array(buf, dtype=np.float32).reshape(n, n).T.copy()


def make_params(p, use_rcp: int = 1, num_threads: int = 1) -> OrcParams:
    """p: bpvo_b200.types.AlgorithmParameters (duck-typed by field name)."""
    c = OrcParams()
    lib().orc_default_params(C.byref(c))
    for name, _ in OrcParams._fields_[:26]:
        v = getattr(p, name)
        setattr(c, name, type(getattr(c, name))(v))
    c.use_rcp, c.num_threads = int(use_rcp), int(num_threads)
    c.dfSigma1, c.dfSigma2 = float(getattr(p, "dfSigma1", 0.75)), float(getattr(p, "dfSigma2", 1.75))
    return c


def _err():
    return RuntimeError(lib().orc_last_error().decode())


# ---- stage functions ---------------------------------------------------------------------------
def pyr_down(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    r, c = img.shape
    out = np.empty(((r + 1) // 2, (c + 1) // 2), dtype=np.uint8)
    lib().orc_pyr_down(_u8(img), r, c, _u8(out))
    return out


def gaussian_blur5(img, sigma):
    img = _f32(img)
    out = np.empty_like(img)
    lib().orc_gaussian_blur5(_fp(img), img.shape[0], img.shape[1], float(sigma), _fp(out))
    return out


def gaussian_blur_f32(img, ksize, sigma):
    """cv::GaussianBlur on CV_32F, any odd ksize (<= 0: from sigma, like cv::Size())"""
    img = _f32(img)
    out = np.empty_like(img)
    lib().orc_gaussian_blur_f32(_fp(img), img.shape[0], img.shape[1], int(ksize), float(sigma), _fp(out))
    return out


def gaussian_blur3_u8(img, sigma):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros_like(img)
    lib().orc_gaussian_blur3_u8(_u8(img), img.shape[0], img.shape[1], float(sigma), _u8(out))
    return out


def census(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    out = np.empty_like(img)
    lib().orc_census(_u8(img), img.shape[0], img.shape[1], _u8(out))
    return out


def descriptor(params, img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    out = np.empty((8,) + img.shape, dtype=np.float32)
    cp = make_params(params)
    n = lib().orc_descriptor(C.byref(cp), _u8(img), img.shape[0], img.shape[1], _fp(out))
    if n < 0:
        raise _err()
    return out[:n].copy()


def saliency(planes):
    planes = _f32(planes)
    c, r, w = planes.shape
    out = np.zeros((r, w), dtype=np.float32)
    lib().orc_saliency(_fp(planes), c, r, w, _fp(out))
    return out


def median(values):
    buf = _f32(values).copy()
    return float(lib().orc_median(_fp(buf), buf.size))


def solve6(H, G):
    Hc, Gc = _colmajor(H), _f32(G).ravel()
    dp = np.zeros(6, dtype=np.float32)
    ok = lib().orc_solve6(_fp(Hc), _fp(Gc), _fp(dp))
    return bool(ok), dp


def params_to_pose(Tn, p):
    out = np.zeros(16, dtype=np.float32)
    lib().orc_params_to_pose(_fp(_colmajor(Tn)), _fp(_f32(p).ravel()), _fp(out))
    return _from_colmajor(out, 4)


# ---- objects -----------------------------------------------------------------------------------
class Frame:
    def __init__(self, K, baseline, rows, cols, params, use_rcp=1, num_threads=1, _borrow=None):
        self._own = _borrow is None
        if _borrow is not None:
            self.h = _borrow
        else:
            cp = make_params(params, use_rcp, num_threads)
            self.h = lib().orc_frame_create(_fp(_colmajor(K)), float(baseline), rows, cols, C.byref(cp))
            if not self.h:
                raise _err()

    def __del__(self):
        if getattr(self, "_own", False) and getattr(self, "h", None):
            lib().orc_frame_destroy(self.h)
            self.h = None

    def set_data(self, img, disp):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        disp = _f32(disp)
        lib().orc_frame_set_data(self.h, _u8(img), _fp(disp))

    def set_template(self):
        if lib().orc_frame_set_template(self.h) != 0:
            raise _err()

    @property
    def num_levels(self):
        return lib().orc_frame_num_levels(self.h)

    def level_size(self, l):
        r, c = C.c_int32(), C.c_int32()
        lib().orc_frame_level_size(self.h, l, C.byref(r), C.byref(c))
        return r.value, c.value

    def pyramid(self, l):
        r, c = self.level_size(l)
        return np.ctypeslib.as_array(lib().orc_frame_pyramid(self.h, l), shape=(r, c)).copy()

    def descriptor(self, l):
        r, c = self.level_size(l)
        ch = C.c_int32()
        p = lib().orc_frame_descriptor(self.h, l, C.byref(ch))
        return np.ctypeslib.as_array(p, shape=(ch.value, r, c)).copy()

    def saliency(self, l):
        r, c = self.level_size(l)
        return np.ctypeslib.as_array(lib().orc_frame_saliency(self.h, l), shape=(r, c)).copy()

    def num_points(self, l):
        return lib().orc_frame_num_points(self.h, l)

    def points(self, l):
        n = self.num_points(l)
        if n == 0:
            return np.zeros((0, 4), np.float32)
        return np.ctypeslib.as_array(lib().orc_frame_points(self.h, l), shape=(n, 4)).copy()

    def num_channels(self, l):
        ch = C.c_int32()
        lib().orc_frame_descriptor(self.h, l, C.byref(ch))
        return ch.value

    def pixels(self, l):
        n, ch = self.num_points(l), self.num_channels(l)
        return np.ctypeslib.as_array(lib().orc_frame_pixels(self.h, l), shape=(ch, n)).copy()

    def jacobians(self, l):
        n, ch = self.num_points(l), self.num_channels(l)
        return np.ctypeslib.as_array(lib().orc_frame_jacobians(self.h, l), shape=(ch * n + 1, 6))[:ch * n].reshape(ch, n, 6).copy()

    def point_inds(self, l):
        n = self.num_points(l)
        return np.ctypeslib.as_array(lib().orc_frame_point_inds(self.h, l), shape=(n,)).copy()

    def normalization(self, l):
        out = np.zeros(16, dtype=np.float32)
        lib().orc_frame_normalization(self.h, l, _fp(out))
        return _from_colmajor(out, 4)


class Estimator:
    def __init__(self, params, num_threads=1, _borrow=None):
        self._own = _borrow is None
        if _borrow is not None:
            self.h = _borrow
        else:
            cp = make_params(params, 1, num_threads)
            self.h = lib().orc_estimator_create(C.byref(cp))

    def __del__(self):
        if getattr(self, "_own", False) and getattr(self, "h", None):
            lib().orc_estimator_destroy(self.h)
            self.h = None

    def linearize(self, ref, cur, level, T, reset_scale=True):
        H = np.zeros(36, np.float32)
        G = np.zeros(6, np.float32)
        s = C.c_float()
        f = lib().orc_linearize(self.h, ref.h, cur.h, level, _fp(_colmajor(T)), int(reset_scale), _fp(H), _fp(G), C.byref(s))
        if f < 0:
            raise _err()
        return dict(f_norm=float(f), H=_from_colmajor(H, 6), G=G.copy(), sigma=float(s.value),
                    residuals=self.residuals(), weights=self.weights(), valid=self.valid())

    def residuals(self):
        n = lib().orc_estimator_num_residuals(self.h)
        return np.ctypeslib.as_array(lib().orc_estimator_residuals(self.h), shape=(n,)).copy()

    def weights(self):
        n = lib().orc_estimator_num_residuals(self.h)
        return np.ctypeslib.as_array(lib().orc_estimator_weights(self.h), shape=(n,)).copy()

    def valid(self):
        n = lib().orc_estimator_num_residuals(self.h)
        return np.ctypeslib.as_array(lib().orc_estimator_valid(self.h), shape=(n,)).copy()

    def estimate_pose(self, ref, cur, T_init):
        L = ref.num_levels
        stats = (OrcStats * L)()
        T = np.zeros(16, np.float32)
        n = lib().orc_estimate_pose(self.h, ref.h, cur.h, _fp(_colmajor(T_init)), _fp(T), stats)
        if n < 0:
            raise _err()
        return _from_colmajor(T, 4), [dict(numIterations=s.numIterations, finalError=s.finalError,
                                          firstOrderOptimality=s.firstOrderOptimality, status=s.status) for s in stats], n

    def fraction_good(self, thresh):
        return float(lib().orc_fraction_good(self.h, float(thresh)))

    def set_trace(self, on=True):
        lib().orc_estimator_set_trace(self.h, int(on))

    def get_trace(self):
        """rows {level, eval, f_norm, |dp|, max|G|, sigma, converged, status} of run() since set_trace(True)"""
        n = lib().orc_estimator_get_trace(self.h, None, 0)
        rows = np.zeros((max(n, 1), 8), np.float32)
        lib().orc_estimator_get_trace(self.h, _fp(rows), n)
        return rows[:n]


class VisualOdometry:
    """restated bpvo::VisualOdometry (bpvo/vo.cc)."""

    def __init__(self, K, baseline, image_size, params, use_rcp=1, num_threads=1):
        rows, cols = image_size
        cp = make_params(params, use_rcp, num_threads)
        self.h = lib().orc_vo_create(_fp(_colmajor(K)), float(baseline), rows, cols, C.byref(cp))
        if not self.h:
            raise _err()
        self.rows, self.cols = rows, cols

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_vo_destroy(self.h)
            self.h = None

    def add_frame(self, img, disp):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        disp = _f32(disp)
        r = OrcResult()
        if lib().orc_vo_add_frame(self.h, _u8(img), _fp(disp), C.byref(r)) != 0:
            raise _err()
        return dict(pose=_from_colmajor(r.pose, 4), isKeyFrame=bool(r.isKeyFrame), keyFramingReason=r.keyFramingReason,
                    stats=[dict(numIterations=s.numIterations, finalError=s.finalError,
                                firstOrderOptimality=s.firstOrderOptimality, status=s.status) for s in r.stats[:r.numLevels]],
                    numFunEvals=r.numFunEvals, numPointCloud=r.numPointCloud)

    def add_frame_raw(self, img_ptr, disp_ptr, result):
        """no-copy call for timing loops: img_ptr/disp_ptr are ctypes pointers, result an OrcResult."""
        return lib().orc_vo_add_frame(self.h, img_ptr, disp_ptr, C.byref(result))

    def num_points_at_level(self, level=-1):
        return lib().orc_vo_num_points_at_level(self.h, level)

    def ref_frame(self):
        return Frame(None, 0, 0, 0, None, _borrow=lib().orc_vo_ref_frame(self.h))

    def estimator(self):
        return Estimator(None, _borrow=lib().orc_vo_estimator(self.h))

    def trajectory(self):
        n = lib().orc_vo_trajectory(self.h, None, 0)
        buf = np.zeros((n, 16), np.float32)
        lib().orc_vo_trajectory(self.h, _fp(buf), n)
        return np.stack([b.reshape(4, 4).T for b in buf]) if n else np.zeros((0, 4, 4), np.float32)

    def point_cloud(self):
        n = lib().orc_vo_point_cloud(self.h, None, None, None, 0)
        xyzw = np.zeros((n, 4), np.float32)
        w = np.zeros(n, np.float32)
        g = np.zeros(n, np.uint8)
        if n:
            lib().orc_vo_point_cloud(self.h, _fp(xyzw), _fp(w), _u8(g), n)
        return xyzw, w, g


# ---- oracle/_ref: the REAL reference code (leaf files) built against header stand-ins; pins the restatement ------
_REF = None


def build_ref() -> str:
    """builds oracle/_ref/libbpvo_ref.so when the reference sources are present (build container); the GPU box
    uses the prebuilt file that travels with the repo snapshot.  Returns the path or '' when unavailable."""
    so = os.path.join(_HERE, "_ref", "libbpvo_ref.so")
    ref_root = os.environ.get("BPVO_REFERENCE", "/root/reference")
    if os.path.isdir(os.path.join(ref_root, "bpvo")):
        deps = [os.path.join(_HERE, f) for f in ("ref_shim.cc", "ref_other_descriptors.cc", "bpvo_oracle.h", os.path.join("refstub", "cv_stub_impl.cc"),
                                                 os.path.join("refstub", "Eigen", "Core"), os.path.join("refstub", "opencv2", "core", "core.hpp"))]
        stale = (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps)
        if stale:
            subprocess.run(["make", "-C", _HERE, "-s", "ref", f"REF={ref_root}"], check=True)
    return so if os.path.exists(so) else ""


def ref_lib():
    global _REF
    if _REF is not None:
        return _REF
    so = build_ref()
    if not so:
        return None
    L = C.CDLL(so)
    fp, u8p, u16p = C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_uint16)
    sig = {
        "ref_census": (None, [u8p, C.c_int, C.c_int, u8p]),
        "ref_census_sigma": (None, [u8p, C.c_int, C.c_int, C.c_float, u8p]),
        "ref_saliency": (None, [fp, C.c_int, C.c_int, C.c_int, fp]),
        "ref_local_max": (None, [fp, C.c_int, C.c_int, C.c_int, C.c_int, u8p]),
        "ref_median": (C.c_float, [fp, C.c_size_t]),
        "ref_compute_weights": (None, [C.c_int, fp, u16p, C.c_size_t, C.c_float, fp]),
        "ref_scale_create": (C.c_void_p, []),
        "ref_scale_destroy": (None, [C.c_void_p]),
        "ref_scale_reset": (None, [C.c_void_p]),
        "ref_scale_estimate": (C.c_float, [C.c_void_p, fp, u16p, C.c_size_t]),
        "ref_linear_system": (C.c_float, [fp, fp, fp, u16p, C.c_size_t, fp, fp]),
        "ref_last_error": (C.c_char_p, []),
        "ref_set_num_threads": (None, [C.c_int]),
        "ref_get_num_threads": (C.c_int, []),
        "ref_frame_create": (C.c_void_p, [fp, C.c_float, C.c_int, C.c_int, C.POINTER(OrcParams)]),
        "ref_frame_destroy": (None, [C.c_void_p]),
        "ref_frame_set_data": (C.c_int, [C.c_void_p, u8p, fp]),
        "ref_frame_set_template": (C.c_int, [C.c_void_p]),
        "ref_frame_num_points": (C.c_int, [C.c_void_p, C.c_int]),
        "ref_frame_level_size": (None, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
        "ref_frame_descriptor": (C.c_int, [C.c_void_p, C.c_int, fp]),
        "ref_frame_points": (None, [C.c_void_p, C.c_int, fp]),
        "ref_frame_pixels": (None, [C.c_void_p, C.c_int, fp]),
        "ref_frame_jacobians": (None, [C.c_void_p, C.c_int, fp]),
        "ref_frame_num_pixels": (C.c_int, [C.c_void_p, C.c_int]),
        "ref_estimator_create": (C.c_void_p, [C.POINTER(OrcParams)]),
        "ref_estimator_destroy": (None, [C.c_void_p]),
        "ref_linearize": (C.c_float, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, fp, C.c_int, fp, fp]),
        "ref_estimator_num_residuals": (C.c_size_t, [C.c_void_p]),
        "ref_estimator_vectors": (None, [C.c_void_p, fp, fp, u16p]),
        "ref_run_level": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, fp, C.POINTER(OrcStats)]),
        "ref_vo_create": (C.c_void_p, [fp, C.c_float, C.c_int, C.c_int, C.POINTER(OrcParams)]),
        "ref_vo_destroy": (None, [C.c_void_p]),
        "ref_vo_add_frame": (C.c_int, [C.c_void_p, u8p, fp, C.POINTER(OrcResult)]),
        "ref_vo_num_points_at_level": (C.c_int, [C.c_void_p, C.c_int]),
        "ref_vo_trajectory": (C.c_int, [C.c_void_p, fp, C.c_int]),
        "ref_vo_point_cloud": (C.c_int, [C.c_void_p, fp, fp, u8p, C.c_int]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _REF = L
    return L


_SEAM = None


def build_gpuseam() -> str:
    """builds oracle/_ref/libbpvo_gpuseam.so -- the reference's own bpvo/vo.cc compiled UNCHANGED against the GPU seam of
    integration/ (INTEGRATION.md Option A) and linked to bpvo_b200/libbpvo_b200.so -- when the reference sources are present;
    the GPU box uses the prebuilt file.  Returns the path or '' when unavailable."""
    so = os.path.join(_HERE, "_ref", "libbpvo_gpuseam.so")
    ref_root = os.environ.get("BPVO_REFERENCE", "/root/reference")
    integ = os.path.join(_HERE, "..", "integration")
    if os.path.isdir(os.path.join(ref_root, "bpvo")) and os.path.exists(os.path.join(_HERE, "..", "bpvo_b200", "libbpvo_b200.so")):
        deps = [os.path.join(integ, f) for f in ("vo_b200_seam.cc", "seam_test_shim.cc", os.path.join("bpvo", "vo_frame.h"),
                                                 os.path.join("bpvo", "vo_pose_estimator.h"))] + [os.path.join(_HERE, "..", "include", "bpvo_b200.h")]
        stale = (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps)
        if stale:
            subprocess.run(["make", "-C", _HERE, "-s", "gpuseam", f"REF={ref_root}"], check=True)
    return so if os.path.exists(so) else ""


def seam_lib():
    """the VisualOdometry-level entry points (same names / layouts as ref_lib()'s) of libbpvo_gpuseam.so"""
    global _SEAM
    if _SEAM is not None:
        return _SEAM
    so = build_gpuseam()
    if not so:
        return None
    L = C.CDLL(so)
    fp, u8p = C.POINTER(C.c_float), C.POINTER(C.c_uint8)
    sig = {
        "ref_last_error": (C.c_char_p, []),
        "ref_vo_create": (C.c_void_p, [fp, C.c_float, C.c_int, C.c_int, C.POINTER(OrcParams)]),
        "ref_vo_destroy": (None, [C.c_void_p]),
        "ref_vo_add_frame": (C.c_int, [C.c_void_p, u8p, fp, C.POINTER(OrcResult)]),
        "ref_vo_num_points_at_level": (C.c_int, [C.c_void_p, C.c_int]),
        "ref_vo_trajectory": (C.c_int, [C.c_void_p, fp, C.c_int]),
        "ref_vo_point_cloud": (C.c_int, [C.c_void_p, fp, fp, u8p, C.c_int]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _SEAM = L
    return L


class RefFrame:
    """the reference's own VisualOdometryFrame (compiled from /root/reference against the stand-in headers)"""

    def __init__(self, K, baseline, rows, cols, params):
        L = ref_lib()
        cp = make_params(params)
        self.L, self.params = L, params
        self.h = L.ref_frame_create(_fp(_colmajor(K)), float(baseline), rows, cols, C.byref(cp))
        if not self.h:
            raise RuntimeError(L.ref_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_frame_destroy(self.h)
            self.h = None

    def set_data(self, img, disp):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        disp = _f32(disp)
        if self.L.ref_frame_set_data(self.h, _u8(img), _fp(disp)) != 0:
            raise RuntimeError(self.L.ref_last_error().decode())

    def set_template(self):
        if self.L.ref_frame_set_template(self.h) != 0:
            raise RuntimeError(self.L.ref_last_error().decode())

    def level_size(self, l):
        r, c = C.c_int32(), C.c_int32()
        self.L.ref_frame_level_size(self.h, l, C.byref(r), C.byref(c))
        return r.value, c.value

    def num_points(self, l):
        return self.L.ref_frame_num_points(self.h, l)

    def descriptor(self, l):
        r, c = self.level_size(l)
        out = np.zeros((8, r, c), np.float32)
        n = self.L.ref_frame_descriptor(self.h, l, _fp(out))
        return out[:n].copy()

    def points(self, l):
        out = np.zeros((self.num_points(l), 4), np.float32)
        self.L.ref_frame_points(self.h, l, _fp(out))
        return out

    def pixels(self, l):
        n, tot = self.num_points(l), self.L.ref_frame_num_pixels(self.h, l)
        out = np.zeros(tot, np.float32)
        self.L.ref_frame_pixels(self.h, l, _fp(out))
        return out.reshape(tot // max(n, 1), n)

    def jacobians(self, l):
        n, tot = self.num_points(l), self.L.ref_frame_num_pixels(self.h, l)
        out = np.zeros((tot, 6), np.float32)
        self.L.ref_frame_jacobians(self.h, l, _fp(out))
        return out.reshape(tot // max(n, 1), n, 6)


class RefEstimator:
    """the reference's own PoseEstimatorGN<TemplateData>"""

    def __init__(self, params):
        self.L = ref_lib()
        cp = make_params(params)
        self.h = self.L.ref_estimator_create(C.byref(cp))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_estimator_destroy(self.h)
            self.h = None

    def linearize(self, ref, cur, level, T, reset=True):
        H = np.zeros(36, np.float32); G = np.zeros(6, np.float32)
        f = self.L.ref_linearize(self.h, ref.h, cur.h, level, _fp(_colmajor(T)), int(reset), _fp(H), _fp(G))
        if f < 0:
            raise RuntimeError(self.L.ref_last_error().decode())
        n = self.L.ref_estimator_num_residuals(self.h)
        r = np.zeros(n, np.float32); w = np.zeros(n, np.float32); v = np.zeros(n, np.uint16)
        self.L.ref_estimator_vectors(self.h, _fp(r), _fp(w), v.ctypes.data_as(C.POINTER(C.c_uint16)))
        return dict(f_norm=float(f), H=_from_colmajor(H, 6), G=G, residuals=r, weights=w, valid=v)

    def run_level(self, ref, cur, level, T):
        Tc = _colmajor(T).copy()
        st = OrcStats()
        if self.L.ref_run_level(self.h, ref.h, cur.h, level, _fp(Tc), C.byref(st)) != 0:
            raise RuntimeError(self.L.ref_last_error().decode())
        return _from_colmajor(Tc, 4), dict(numIterations=st.numIterations, finalError=st.finalError,
                                           firstOrderOptimality=st.firstOrderOptimality, status=st.status)


class RefVisualOdometry:
    """the reference's own bpvo::VisualOdometry (lib = seam_lib(): its vo.cc on top of the GPU seam instead of its CPU classes)"""

    def __init__(self, K, baseline, image_size, params, lib=None):
        self.L = lib if lib is not None else ref_lib()
        rows, cols = image_size
        cp = make_params(params)
        self.h = self.L.ref_vo_create(_fp(_colmajor(K)), float(baseline), rows, cols, C.byref(cp))
        if not self.h:
            raise RuntimeError(self.L.ref_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_vo_destroy(self.h)
            self.h = None

    def add_frame(self, img, disp):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        disp = _f32(disp)
        r = OrcResult()
        if self.L.ref_vo_add_frame(self.h, _u8(img), _fp(disp), C.byref(r)) != 0:
            raise RuntimeError(self.L.ref_last_error().decode())
        return dict(pose=_from_colmajor(r.pose, 4), isKeyFrame=bool(r.isKeyFrame), keyFramingReason=r.keyFramingReason,
                    stats=[dict(numIterations=s.numIterations, finalError=s.finalError, firstOrderOptimality=s.firstOrderOptimality,
                                status=s.status) for s in r.stats[:r.numLevels]], numPointCloud=r.numPointCloud)

    def add_frame_raw(self, img_ptr, disp_ptr, result):
        return self.L.ref_vo_add_frame(self.h, img_ptr, disp_ptr, C.byref(result))

    def num_points_at_level(self, level=-1):
        return self.L.ref_vo_num_points_at_level(self.h, level)

    def trajectory(self):
        n = self.L.ref_vo_trajectory(self.h, None, 0)
        buf = np.zeros((max(n, 1), 16), np.float32)
        self.L.ref_vo_trajectory(self.h, _fp(buf), n)
        return np.stack([b.reshape(4, 4).T for b in buf[:n]]) if n else np.zeros((0, 4, 4), np.float32)

    def point_cloud(self, n):
        xyzw = np.zeros((n, 4), np.float32); w = np.zeros(n, np.float32); g = np.zeros(n, np.uint8)
        self.L.ref_vo_point_cloud(self.h, _fp(xyzw), _fp(w), _u8(g), n)
        return xyzw, w, g


def stereo_prefilter_xsobel(img, cap=31):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    lib().orc_stereo_prefilter_xsobel(_u8(img), img.shape[0], img.shape[1], int(cap), _u8(out))
    return out


def stereo_bm(left, right, numberOfDisparities, SADWindowSize=15, minDisparity=0, preFilterCap=31, textureThreshold=10,
              uniquenessRatio=15):
    """OpenCV StereoBM as the reference configures it (utils/stereo_algorithm.cc:67-111) -> (int16 fixed point, float32 = /16)."""
    left = np.ascontiguousarray(left, np.uint8); right = np.ascontiguousarray(right, np.uint8)
    assert left.shape == right.shape and left.ndim == 2
    sp = OrcStereoParams(int(numberOfDisparities), int(SADWindowSize), int(minDisparity), int(preFilterCap), int(textureThreshold),
                         int(uniquenessRatio))
    d16 = np.empty(left.shape, np.int16); df = np.empty(left.shape, np.float32)
    rc = lib().orc_stereo_bm(_u8(left), _u8(right), left.shape[0], left.shape[1], C.byref(sp), d16.ctypes.data_as(C.POINTER(C.c_int16)), _fp(df))
    if rc != 0:
        raise ValueError("orc_stereo_bm: unsupported parameters")
    return d16, df
