"""CPU tests of the drop-in boundary: the C-ABI library loads, exports exactly what include/bpvo_b200.h declares,
mirrors the reference's defaults, and fails LOUDLY (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import HAS_GPU, ROOT, make_params


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "bpvo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bpvo_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from bpvo_b200 import _capi
    lib = _capi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 45
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/bpvo_b200.h but not exported"
    assert sorted(_capi.SIGNATURES) == declared, "python binding and header disagree"
    assert lib.bpvo_b200_version() == 100


def test_default_params_are_the_reference_defaults():
    from bpvo_b200 import _capi
    from bpvo_b200.types import AlgorithmParameters, CParams, fill_cparams
    c = CParams()
    _capi.lib().bpvo_b200_default_params(C.byref(c))
    py = fill_cparams(AlgorithmParameters())
    for name, _ in CParams._fields_[:26]:
        assert getattr(c, name) == getattr(py, name), name
    # bpvo/types.cc:31-66
    assert (c.numPyramidLevels, c.sigmaBitPlanes, c.maxIterations, c.lossFunction, c.descriptor) == (-1, 0.5, 50, 0x11, 0x30)
    assert (c.minNumPixelsForNonMaximaSuppression, c.nonMaxSuppRadius, c.withNormalization) == (76800, 1, 1)
    assert abs(c.parameterTolerance - 1e-7) < 1e-12 and abs(c.minSaliency - 0.1) < 1e-7
    assert C.sizeof(CParams) == 30 * 4      # 26 reference fields + device_id, flags + dfSigma1, dfSigma2


def test_auto_pyramid_levels_rule():
    from bpvo_b200.types import AlgorithmParameters
    p = AlgorithmParameters()
    assert p.resolved_num_levels(480, 640) == 5        # 1 + round(log2(480/40)) (vo.cc:101-104)
    assert p.resolved_num_levels(376, 1241) == 4
    assert p.resolved_num_levels(1080, 1920) == 6


@pytest.mark.skipif(HAS_GPU, reason="checks the no-device behaviour")
def test_no_cpu_fallback():
    from bpvo_b200 import Error, VisualOdometry, _capi
    from bpvo_b200.engine import Context
    assert _capi.lib().bpvo_b200_device_count() == 0
    with pytest.raises(Error, match="no CUDA device"):
        VisualOdometry(np.eye(3), 0.1, (96, 128), make_params("intensity", 2))
    with pytest.raises(Error, match="no CUDA device"):
        Context(np.eye(3), 0.1, (96, 128), make_params("bitplanes", 2))


def test_argument_validation_happens_before_touching_the_device():
    from bpvo_b200 import Error
    from bpvo_b200.engine import Context
    from bpvo_b200.types import DescriptorType
    with pytest.raises(Error, match="pyramid"):
        Context(np.eye(3), 0.1, (96, 128), make_params("intensity", 40))
    bad = make_params("intensity", 2); bad.descriptor = DescriptorType.kLatch
    with pytest.raises(Error, match="not on the accelerated path"):
        Context(np.eye(3), 0.1, (96, 128), bad)
    K = np.eye(3); K[0, 1] = 0.5
    with pytest.raises(Error, match="K must be"):
        Context(K, 0.1, (96, 128), make_params("intensity", 2))


def test_product_does_not_import_the_oracle():
    """the oracle is test infrastructure: nothing under bpvo_b200/ may reference it"""
    pkg = os.path.join(ROOT, "bpvo_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(root, f), errors="ignore").read()
                assert "pyoracle" not in src and "liboracle" not in src and "bpvo_oracle" not in src, os.path.join(root, f)


def test_synthetic_scene_is_deterministic():
    from bpvo_b200 import synth
    a = synth.scene_small(48, 64).render(1)
    b = synth.scene_small(48, 64).render(1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert a[0].dtype == np.uint8 and a[1].dtype == np.float32 and a[1].min() > 0
    sc = synth.scene_small(48, 64, hole_fraction=0.2)
    assert 0.1 < (sc.render(0)[1] == 0).mean() < 0.3
    T = sc.relative_pose(0, 1)
    assert np.abs(T[:3, :3] @ T[:3, :3].T - np.eye(3)).max() < 1e-12


def test_cpp_example_links_against_the_host_shim(tmp_path):
    """examples/vo_stream.cpp uses bpvo_b200::VisualOdometry (the C++ mirror of bpvo/vo.h); it must compile
    and link against the in-tree library with a plain g++ (no compute here: there is no GPU)."""
    import subprocess
    libdir = os.path.join(ROOT, "bpvo_b200")
    exe = tmp_path / "vo_stream"
    subprocess.run(["/usr/bin/g++", "-std=c++14", "-O2", os.path.join(ROOT, "examples", "vo_stream.cpp"),
                    "-I" + os.path.join(libdir, "csrc", "host"), "-L" + libdir, "-lbpvo_b200", "-Wl,-rpath," + libdir,
                    "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


def test_shared_memory_plan_of_the_device_loop():
    """tpl_cache_plan (kernels_linearize.cuh) evaluated on the host: whole fields in the order residuals + valid, points,
    gx, gy, I0; no overlap, inside the budget, 16-byte aligned points; the three regimes DESIGN.md names"""
    import ctypes as C
    from bpvo_b200 import _capi
    lib = _capi.lib()
    NONE = 0xffffffff
    budget = (227 * 1024 - 24 * 1024 - 5120) // 1024 * 1024          # what engine.cu requests on a B200

    def plan(ch, bytes_, need):
        out = (C.c_uint32 * 8)()
        assert lib.bpvo_b200_debug_cache_plan(ch, bytes_, need, out) == 0
        return dict(pts=out[0], i0=out[1], gx=out[2], gy=out[3], r=out[4], valid=out[5], K=out[6], first=out[7])

    def check(ch, bytes_, need):
        p = plan(ch, bytes_, need)
        size = dict(pts=need * 256 * 16, i0=need * 256 * 4 * ch, gx=need * 256 * 4 * ch, gy=need * 256 * 4 * ch, r=need * 256 * 4 * ch, valid=need * 256)
        spans = sorted((p[k], p[k] + size[k]) for k in size if p[k] != NONE)
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 <= b0, (p, "overlap")
        if spans:
            assert spans[0][0] >= p["first"] and spans[-1][1] <= p["first"] + bytes_, (p, "outside the cache area")
        assert (p["r"] == NONE) == (p["valid"] == NONE)
        if p["pts"] != NONE:
            assert p["pts"] % 16 == 0
        # priority: a field is only cached when everything before it in the order is
        order = ["r", "pts", "gx", "gy", "i0"]
        cached = [p[k] != NONE for k in order]
        # (all-or-nothing per field, greedy: a later, smaller field may still fit when an earlier one did not -- only r -> pts differ in size)
        assert cached[2] >= cached[3] >= cached[4]
        return p

    semi = check(8, budget, 1)                        # KITTI semi-dense: everything on chip
    assert all(semi[k] != NONE for k in ("pts", "i0", "gx", "gy", "r", "valid"))
    dense = check(8, budget, 11)                      # KITTI dense level 0: residuals + points
    assert dense["r"] != NONE and dense["pts"] != NONE and dense["gx"] == NONE and dense["i0"] == NONE
    shard = check(8, budget, 6)                       # a 1080p level over 8 GPUs with 6 points per thread: all but I0
    assert shard["gy"] != NONE and shard["i0"] == NONE
    shard7 = check(8, budget, 7)                      # ... and with 7 (228 576 points on 148 x 256 threads): still all but I0
    assert shard7["gy"] != NONE and shard7["i0"] == NONE
    huge = check(8, budget, 48)                       # dense 1080p level 0 on one GPU: only the 3-D points fit
    assert huge["pts"] != NONE and all(huge[k] == NONE for k in ("i0", "gx", "gy", "r"))
    assert all(check(8, budget, 60)[k] == NONE for k in ("pts", "i0", "gx", "gy", "r"))
    for ch in (1, 8):
        for need in range(0, 60):
            for b in (0, 4096, 50000, budget):
                check(ch, b, need)


def test_stereo_params_mirror_the_reference_defaults():
    """bpvo_b200_stereo_params = the CvStereoBMState fields utils/stereo_algorithm.cc:67-85 sets, with OpenCV's defaults"""
    from bpvo_b200 import _capi
    p = _capi.CStereoParams()
    assert C.sizeof(p) == 13 * 4
    _capi.lib().bpvo_b200_stereo_default_params(C.byref(p))
    assert (p.preFilterType, p.preFilterSize, p.preFilterCap, p.SADWindowSize, p.minDisparity, p.numberOfDisparities) == (1, 9, 31, 15, 0, 0)
    assert (p.textureThreshold, p.uniquenessRatio, p.speckleWindowSize, p.speckleRange, p.trySmallerWindows, p.disp12MaxDiff) == (10, 15, 0, 0, 0, -1)


@pytest.mark.skipif(HAS_GPU, reason="checks the no-device behaviour")
def test_stereo_create_fails_loudly_without_a_device():
    from bpvo_b200 import Error
    from bpvo_b200.stereo import StereoAlgorithm
    with pytest.raises(Error, match="no CUDA device"):
        StereoAlgorithm((64, 128), numberOfDisparities=16)
    with pytest.raises(Error, match="numberOfDisparities"):
        StereoAlgorithm((64, 128))                                        # "must be provided" (stereo_algorithm.cc:75)
    with pytest.raises(Error, match="divisble by 16"):
        StereoAlgorithm((64, 128), numberOfDisparities=20)                # OpenCV's own check and message
    with pytest.raises(Error, match="accelerated path"):
        StereoAlgorithm((64, 128), numberOfDisparities=16, StereoAlgorithm="SGBM")
