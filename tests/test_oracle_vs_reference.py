"""Pins the oracle against the REAL reference code: oracle/_ref/libbpvo_ref.so is compiled from the reference's own
mestimator.cc, census.cc, imgproc.cc, linear_system_builder.cc, utils.cc (+ header-only IsLocalMax / median) where they
lie under /root/reference, against small header stand-ins for Eigen/OpenCV (oracle/refstub).  These are exactly the
quirk-laden SIMD pieces (saliency store bug, 3x4 NMS window, weights ignoring `valid`, packed rank-1 update, size_t
arithmetic in the scale).  Everything here must match BIT FOR BIT.  Runs on CPU; skipped if the library is absent."""
import ctypes as C

import numpy as np
import pytest

from conftest import make_params


@pytest.fixture(scope="module")
def ref(oracle):
    L = oracle.ref_lib()
    if L is None:
        pytest.skip("oracle/_ref/libbpvo_ref.so not available (reference sources absent and no prebuilt library)")
    return L


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _u16(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint16))


@pytest.mark.parametrize("shape", [(20, 33), (37, 64), (40, 18), (94, 311)])
def test_census_bit_exact(oracle, ref, shape):
    rng = np.random.RandomState(shape[1])
    img = rng.randint(0, 256, size=shape).astype(np.uint8)
    img[5:9, 4:12] = 100
    out = np.zeros_like(img)
    ref.ref_census(_u8(img), shape[0], shape[1], _u8(out))
    assert np.array_equal(oracle.census(img), out)


@pytest.mark.parametrize("shape,sigma", [((20, 33), 0.75), ((37, 64), 0.5), ((94, 311), 0.75), ((40, 18), 1.3)])
def test_census_with_pre_blur_bit_exact(oracle, ref, shape, sigma):
    """sigmaPriorToCensusTransform > 0 (census.cc:63-65): the reference's census() over the stand-in cv::GaussianBlur(3x3, u8)
    == the oracle's fixed-point blur (pinned against cv2 golden vectors in test_oracle_cpu.py) followed by its census"""
    rng = np.random.RandomState(shape[0])
    img = rng.randint(0, 256, size=shape).astype(np.uint8)
    out = np.zeros_like(img)
    ref.ref_census_sigma(_u8(img), shape[0], shape[1], sigma, _u8(out))
    assert np.array_equal(oracle.census(oracle.gaussian_blur3_u8(img, sigma)), out)


@pytest.mark.parametrize("shape", [(12, 16), (11, 19), (9, 22), (47, 156), (94, 311), (60, 80)])
@pytest.mark.parametrize("channels", [1, 8])
def test_saliency_bit_exact_including_the_bugs(oracle, ref, shape, channels):
    rng = np.random.RandomState(7 * channels + shape[1])
    planes = rng.rand(channels, *shape).astype(np.float32)
    out = np.zeros(shape, np.float32)
    ref.ref_saliency(_fp(planes), channels, shape[0], shape[1], _fp(out))
    assert np.array_equal(oracle.saliency(planes), out)


def test_saliency_on_a_real_descriptor(oracle, ref):
    from bpvo_b200 import synth
    img, _ = synth.scene_small(96, 128).render(0)
    planes = oracle.descriptor(make_params("bitplanes", 1), img)
    out = np.zeros(img.shape, np.float32)
    ref.ref_saliency(_fp(planes), 8, img.shape[0], img.shape[1], _fp(out))
    mine = oracle.saliency(planes)
    assert np.array_equal(mine, out)
    assert mine[:, 4:-4].max() <= 2.0          # columns >= 4 only ever see channel 0 (store-to-dst bug, Q3)


@pytest.mark.parametrize("radius", [1, 2])
def test_selection_matches_reference_local_max(oracle, ref, radius):
    """TemplateData::setData's candidate scan with the reference's IsLocalMax vs the oracle's selected pixel list"""
    from bpvo_b200 import synth
    sc = synth.scene_small(96, 128)
    img, d = sc.render(0)
    p = make_params("intensity", 1, nonMaxSuppRadius=radius, minNumPixelsForNonMaximaSuppression=1, minSaliency=1.0)
    f = oracle.Frame(sc.K, sc.baseline, sc.rows, sc.cols, p)
    f.set_data(img, d); f.set_template()
    S = f.saliency(0)
    flags = np.zeros(S.shape, np.uint8)
    border = max(radius, 3)
    ref.ref_local_max(_fp(S), S.shape[0], S.shape[1], radius, border, _u8(flags))
    expect = np.flatnonzero((flags.ravel() == 1) & (S.ravel() >= p.minSaliency))
    expect = expect[: len(expect) // 16 * 16]
    assert np.array_equal(f.point_inds(0), expect.astype(np.int32))


def test_median_bit_exact(oracle, ref):
    rng = np.random.RandomState(11)
    for n in [1, 2, 3, 4, 5, 16, 101, 1000, 4097]:
        v = rng.rand(n).astype(np.float32)
        assert oracle.median(v) == ref.ref_median(_fp(v), n)


@pytest.mark.parametrize("loss", [0x10, 0x11, 0x12])
def test_weights_bit_exact(oracle, ref, loss):
    """through the oracle's linearize (weights are an output of it) vs MEstimator::ComputeWeights on the same r, valid, sigma"""
    from bpvo_b200 import synth
    sc = synth.scene_small(96, 128)
    name = {0x10: "huber", 0x11: "tukey", 0x12: "l2"}[loss]
    p = make_params("bitplanes", 2, name)
    a = oracle.Frame(sc.K, sc.baseline, sc.rows, sc.cols, p); b = oracle.Frame(sc.K, sc.baseline, sc.rows, sc.cols, p)
    i0, d0 = sc.render(0); i1, d1 = sc.render(1)
    a.set_data(i0, d0); a.set_template(); b.set_data(i1, d1)
    T = np.eye(4, dtype=np.float32); T[0, 3] = 0.3          # pushes some points out: invalid entries keep weight f(0) (Q6)
    o = oracle.Estimator(p).linearize(a, b, 0, T)
    r, v = np.ascontiguousarray(o["residuals"]), np.ascontiguousarray(o["valid"])
    assert 0 < v.sum() < v.size
    w = np.zeros_like(r)
    ref.ref_compute_weights(loss, _fp(r), _u16(v), r.size, C.c_float(o["sigma"]), _fp(w))
    assert np.array_equal(w, o["weights"])
    # scale: AutoScaleEstimator::estimateScale on the same vectors, fresh state
    h = ref.ref_scale_create()
    s = ref.ref_scale_estimate(h, _fp(r), _u16(v), r.size)
    ref.ref_scale_destroy(h)
    assert s == o["sigma"]
    # normal equations: LinearSystemBuilder::Run on the oracle's Jacobians / residuals / weights
    J = np.ascontiguousarray(a.jacobians(0).reshape(-1, 6))
    H = np.zeros(36, np.float32); G = np.zeros(6, np.float32)
    f = ref.ref_linear_system(_fp(J), _fp(r), _fp(w), _u16(v), r.size, _fp(H), _fp(G))
    assert f == o["f_norm"]
    assert np.array_equal(H.reshape(6, 6).T, o["H"]) and np.array_equal(G, o["G"])


def test_scale_estimator_state_machine(oracle, ref):
    """delta-scale gate across successive calls (mestimator.cc:467-490): same sequence of sigmas"""
    from bpvo_b200 import synth
    sc = synth.scene_small(96, 128)
    p = make_params("intensity", 2, "huber")
    a = oracle.Frame(sc.K, sc.baseline, sc.rows, sc.cols, p); b = oracle.Frame(sc.K, sc.baseline, sc.rows, sc.cols, p)
    i0, d0 = sc.render(0); i1, d1 = sc.render(1)
    a.set_data(i0, d0); a.set_template(); b.set_data(i1, d1)
    est = oracle.Estimator(p)
    h = ref.ref_scale_create()
    T = np.eye(4, dtype=np.float32)
    for it in range(4):
        o = est.linearize(a, b, 1, T, reset_scale=(it == 0))
        r, v = np.ascontiguousarray(o["residuals"]), np.ascontiguousarray(o["valid"])
        if it == 0:
            ref.ref_scale_reset(h)
        s = ref.ref_scale_estimate(h, _fp(r), _u16(v), r.size)
        assert s == o["sigma"], it
        if it == 1:
            T = T.copy(); T[0, 3] = 0.01
    ref.ref_scale_destroy(h)


# =====================================================================================================================
# The WHOLE hot path of the real reference (vo.cc, vo_frame.cc, vo_pose_estimator.cc, pose_estimator_{base,gn}.h,
# template_data.cc, rigid_body_warp.cc incl. its _mm_rcp_ps Jacobians, warps.cc, photo_error.cc, the descriptors, ...)
# compiled from /root/reference against the stand-in Eigen/OpenCV headers: the oracle must reproduce it BIT FOR BIT,
# down to the iteration counts of every pyramid level and the key-frame decisions.
# =====================================================================================================================
def _frames(kind):
    from bpvo_b200 import synth
    if kind == "small":
        return synth.scene_small(96, 128)
    if kind == "odd":
        return synth.scene_small(117, 203, seed=11)
    return synth.scene_vga()


@pytest.mark.parametrize("kind,desc,levels,loss", [("small", "bitplanes", 3, "tukey"), ("small", "intensity", 3, "huber"),
                                                    ("odd", "bitplanes", 2, "huber"), ("odd", "intensity", 2, "l2"),
                                                    ("vga", "intensity", 4, "huber"),
                                                    # the reference's own bpvo/gradient_descriptor.cc: GradientDescriptor (3 channels,
                                                    # without / with the Gaussian of cv::Size() size), DescriptorFields (5 channels)
                                                    ("small", "gradient", 3, "huber"), ("odd", "gradient:0.75", 2, "tukey"), ("odd", "gradient:1.2", 2, "tukey"),
                                                    ("small", "dfields", 3, "tukey"), ("odd", "dfields", 2, "huber")])
def test_frame_and_linearize_bit_exact(oracle, ref, kind, desc, levels, loss):
    sc = _frames(kind)
    desc, _, sig = desc.partition(":")
    p = make_params(desc, levels, loss, **({"sigmaPriorToCensusTransform": float(sig)} if sig else {}))
    i0, d0 = sc.render(0); i1, d1 = sc.render(1)
    ra, rb = oracle.RefFrame(sc.K, sc.baseline, sc.rows, sc.cols, p), oracle.RefFrame(sc.K, sc.baseline, sc.rows, sc.cols, p)
    oa, ob = oracle.Frame(sc.K, sc.baseline, sc.rows, sc.cols, p, use_rcp=1), oracle.Frame(sc.K, sc.baseline, sc.rows, sc.cols, p, use_rcp=1)
    for a, b in ((ra, rb), (oa, ob)):
        a.set_data(i0, d0); a.set_template(); b.set_data(i1, d1)
    re, oe = oracle.RefEstimator(p), oracle.Estimator(p)
    T1 = np.array(sc.relative_pose(0, 1), dtype=np.float32)
    for l in range(levels):
        assert ra.num_points(l) == oa.num_points(l) > 0
        assert np.array_equal(ra.descriptor(l), oa.descriptor(l))
        assert np.array_equal(ra.points(l), oa.points(l))                 # makePoint
        assert np.array_equal(ra.pixels(l), oa.pixels(l))
        assert np.array_equal(ra.jacobians(l), oa.jacobians(l))           # SSE Jacobians with _mm_rcp_ps, Hartley normalisation
        for it, T in enumerate([np.eye(4, dtype=np.float32), T1]):
            r = re.linearize(ra, rb, l, T, reset=(it == 0))
            o = oe.linearize(oa, ob, l, T, reset_scale=(it == 0))
            for key in ("residuals", "valid", "weights", "H", "G"):
                assert np.array_equal(r[key], o[key]), (l, it, key)
            assert r["f_norm"] == o["f_norm"]


@pytest.mark.parametrize("desc,opts", [("intensity", dict(interp=1)), ("bitplanes", dict(interp=1)),          # kCosine
                                       ("intensity", dict(interp=2)), ("bitplanes", dict(interp=2)),          # kCubic
                                       ("intensity", dict(interp=3)), ("bitplanes", dict(interp=3)),          # kCubicHermite
                                       ("bitplanes", dict(gradientEstimation=1)),                             # kCentralDifference_5
                                       ("bitplanes", dict(sigmaPriorToCensusTransform=0.75, sigmaBitPlanes=1.618)),
                                       ("intensity", dict(nonMaxSuppRadius=2)), ("bitplanes", dict(withNormalization=0))])
def test_option_variants_bit_exact(oracle, ref, desc, opts):
    """the options beside the defaults -- the other InterpolationTypes (photo_error.cc:391-444), the 5-point gradient
    (template_data.cc:125-129), the pre-census blur (census.cc:63-65), NMS radius 2, no Hartley normalisation -- through
    the reference's own sources vs the oracle: descriptors, template, residuals, weights, H, G bit for bit.
    (Cubic / Hermite: points whose footprint reaches the row past the image are excluded -- the reference reads whatever
    follows its buffer there, photo_error.cc:358.)"""
    sc = _frames("small")
    levels = 2
    p = make_params(desc, levels, "tukey", **opts)
    i0, d0 = sc.render(0); i1, d1 = sc.render(1)
    ra, rb = oracle.RefFrame(sc.K, sc.baseline, sc.rows, sc.cols, p), oracle.RefFrame(sc.K, sc.baseline, sc.rows, sc.cols, p)
    oa, ob = oracle.Frame(sc.K, sc.baseline, sc.rows, sc.cols, p, use_rcp=1), oracle.Frame(sc.K, sc.baseline, sc.rows, sc.cols, p, use_rcp=1)
    for a, b in ((ra, rb), (oa, ob)):
        a.set_data(i0, d0); a.set_template(); b.set_data(i1, d1)
    re, oe = oracle.RefEstimator(p), oracle.Estimator(p)
    T1 = np.array(sc.relative_pose(0, 1), dtype=np.float32)
    footprint4 = opts.get("interp", 0) in (2, 3)
    for l in range(levels):
        N = ra.num_points(l)
        assert N == oa.num_points(l) > 0
        assert np.array_equal(ra.descriptor(l), oa.descriptor(l))
        assert np.array_equal(ra.points(l), oa.points(l))
        assert np.array_equal(ra.pixels(l), oa.pixels(l))
        assert np.array_equal(ra.jacobians(l), oa.jacobians(l))
        r = re.linearize(ra, rb, l, T1, reset=True)
        o = oe.linearize(oa, ob, l, T1, reset_scale=True)
        assert np.array_equal(r["valid"], o["valid"])
        if not footprint4:
            for key in ("residuals", "weights", "H", "G"):
                assert np.array_equal(r[key], o[key]), (l, key)
        else:
            rows = sc.rows >> l if sc.rows % (1 << l) == 0 else (sc.rows + (1 << l) - 1) >> l
            K = np.array(sc.K, np.float64) / (1 << l); K[2, 2] = 1.0
            X = ra.points(l).astype(np.float64).reshape(-1, 4)
            h = (K.astype(np.float32).astype(np.float64) @ T1[:3, :].astype(np.float64)) @ X.T
            keep = np.floor(h[1] / h[2]) < rows - 2.5
            C = r["residuals"].size // N
            assert keep.sum() > 0.9 * N
            assert np.array_equal(r["residuals"].reshape(C, N)[:, keep], o["residuals"].reshape(C, N)[:, keep]), l


@pytest.mark.parametrize("desc,levels,loss,nframes", [("bitplanes", 3, "tukey", 14), ("intensity", 3, "huber", 14), ("intensity", 2, "l2", 6)])
def test_vo_stream_bit_exact(oracle, ref, desc, levels, loss, nframes):
    """addFrame state machine incl. key-framing + re-estimation, trajectory and point cloud: oracle == real reference"""
    sc = _frames("small")
    p = make_params(desc, levels, loss)
    vr = oracle.RefVisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p)
    vo = oracle.VisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p, use_rcp=1)
    n_kf = 0
    for k in range(nframes):
        img, d = sc.render(k)
        a, b = vr.add_frame(img, d), vo.add_frame(img, d)
        assert a["isKeyFrame"] == b["isKeyFrame"] and a["keyFramingReason"] == b["keyFramingReason"], k
        assert np.array_equal(a["pose"], b["pose"]), k
        assert [s["numIterations"] for s in a["stats"]] == [s["numIterations"] for s in b["stats"]], k
        assert [s["status"] for s in a["stats"]] == [s["status"] for s in b["stats"]], k
        assert np.array_equal(np.float32([s["finalError"] for s in a["stats"]]), np.float32([s["finalError"] for s in b["stats"]]))
        assert vr.num_points_at_level() == vo.num_points_at_level()
        assert a["numPointCloud"] == b["numPointCloud"]
        if k > 0 and a["isKeyFrame"]:
            n_kf += 1
            xr, wr, gr = vr.point_cloud(a["numPointCloud"])
            xo, wo, go = vo.point_cloud()
            assert np.array_equal(xr, xo) and np.array_equal(wr, wo) and np.array_equal(gr, go)
    assert np.array_equal(vr.trajectory(), vo.trajectory())
    if nframes >= 14:
        assert n_kf >= 1, "the stream should trigger at least one key-frame"


def test_kitti_pair_bit_exact(oracle, ref):
    """one full-size KITTI bit-planes pair through the reference's own estimatePose path"""
    from bpvo_b200 import synth
    sc = synth.scene_kitti()
    p = make_params("bitplanes", 4, "tukey")
    vr = oracle.RefVisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p)
    vo = oracle.VisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p, use_rcp=1)
    for k in range(2):
        img, d = sc.render(k)
        a, b = vr.add_frame(img, d), vo.add_frame(img, d)
    assert np.array_equal(a["pose"], b["pose"])
    assert [s["numIterations"] for s in a["stats"]] == [s["numIterations"] for s in b["stats"]]
    assert vr.num_points_at_level(0) == vo.num_points_at_level(0)


def test_standin_image_ops_match_cv2_golden(oracle, ref):
    """the stand-in cv::pyrDown / cv::GaussianBlur behind oracle/_ref are the formulas pinned against cv2 4.13 (via the
    descriptors of a frame: pyramid levels and blurred bit-planes must equal the oracle's, which is checked against cv2)"""
    from bpvo_b200 import synth
    sc = synth.scene_small(94, 131)
    p = make_params("bitplanes", 3, sigmaBitPlanes=1.618)
    img, d = sc.render(0)
    r = oracle.RefFrame(sc.K, sc.baseline, sc.rows, sc.cols, p); o = oracle.Frame(sc.K, sc.baseline, sc.rows, sc.cols, p)
    r.set_data(img, d); o.set_data(img, d)
    for l in range(3):
        assert np.array_equal(r.descriptor(l), o.descriptor(l))
