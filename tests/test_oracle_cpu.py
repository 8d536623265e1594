"""CPU tests of the oracle (the checker itself): third-party image ops against cv2 golden vectors,
independent numpy restatements of the quirky pieces, solver algebra, and end-to-end behaviour on the
synthetic scenes.  No GPU needed."""
import os

import numpy as np
import pytest

from conftest import make_params

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_pyrdown_matches_cv2_golden(oracle):
    g = np.load(os.path.join(GOLD, "cv2_golden.npz"))
    for i in range(4):
        out = oracle.pyr_down(g[f"pyr_in_{i}"])
        assert np.array_equal(out, g[f"pyr_out_{i}"]), f"cv::pyrDown case {i} not bit-exact"


def test_gaussian_blur_matches_cv2_golden(oracle):
    g = np.load(os.path.join(GOLD, "cv2_golden.npz"))
    for i in ["0", "1", "2", "3", "f"]:
        out = oracle.gaussian_blur5(g[f"blur_in_{i}"], float(g[f"blur_sigma_{i}"]))
        ref = g[f"blur_out_{i}"]
        assert np.abs(out - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max()), f"cv::GaussianBlur case {i}"


def _census_numpy(img):
    r, c = img.shape
    out = np.zeros_like(img)
    ctr = img[1:-1, 1:-1]
    k = 0
    acc = np.zeros((r - 2, c - 2), np.uint8)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dy == 0 and dx == 0:
                continue
            nb = img[1 + dy:r - 1 + dy, 1 + dx:c - 1 + dx]
            acc |= ((nb >= ctr).astype(np.uint8) << k)
            k += 1
    out[1:-1, 1:-1] = acc
    return out


def test_gaussian_blur3_u8_matches_cv2_golden(oracle):
    """cv::GaussianBlur(3x3) on CV_8U before the census (census.cc:65): bit-exact against cv2 4.13"""
    g = np.load(os.path.join(GOLD, "cv2_golden.npz"))
    n = 0
    while f"cblur_in_{n}" in g.files:
        out = oracle.gaussian_blur3_u8(g[f"cblur_in_{n}"], float(g[f"cblur_sigma_{n}"]))
        assert np.array_equal(out, g[f"cblur_out_{n}"]), n
        n += 1
    assert n >= 5


def test_census_against_numpy(oracle):
    rng = np.random.RandomState(1)
    for shape in [(20, 33), (37, 64), (19, 18)]:
        img = rng.randint(0, 256, size=shape).astype(np.uint8)
        img[3:6, 3:9] = 77          # ties: >= must set the bit
        assert np.array_equal(oracle.census(img), _census_numpy(img))


def _saliency_closed_form(planes):
    """independent statement of what the reference's two saliency functions leave in memory"""
    C, R, W = planes.shape
    flat = planes.reshape(C, -1).astype(np.float32)

    def gm(c, p):
        return np.float32(abs(flat[c, p - 1] - flat[c, p + 1])) + np.float32(abs(flat[c, p - W] - flat[c, p + W]))

    def tail(c, p):
        return np.float32(abs(flat[c, p + 1] - flat[c, p - 1])) + np.float32(abs(flat[c, p + W] + flat[c, p - W]))

    n = W & ~3
    S = np.zeros((R, W), np.float32)
    for y in range(1, R - 1):
        for x in range(W - 1):
            p = y * W + x
            if C == 1:
                S[y, x] = gm(0, p) if x < n else tail(0, p)
            elif x >= n:
                v = tail(0, p)
                for c in range(1, C):
                    v = np.float32(v + tail(c, p))
                S[y, x] = v
            elif x >= 4:
                S[y, x] = gm(0, p)
            else:
                j = n - 4 + x
                pj = y * W + j
                base = np.float32(0) if j == W - 1 else gm(0, pj)
                S[y, x] = np.float32(base + gm(C - 1, pj))
    return S


@pytest.mark.parametrize("shape", [(12, 16), (11, 19), (9, 22)])
@pytest.mark.parametrize("channels", [1, 8])
def test_saliency_literal_equals_closed_form(oracle, shape, channels):
    rng = np.random.RandomState(3)
    planes = rng.rand(channels, *shape).astype(np.float32)
    assert np.array_equal(oracle.saliency(planes), _saliency_closed_form(planes))


def test_median_rule(oracle):
    rng = np.random.RandomState(5)
    for n in [3, 4, 5, 8, 101, 1000]:
        v = rng.rand(n).astype(np.float32)
        s = np.sort(v)
        expect = s[n // 2] if n % 2 else np.float32((np.float32(s[n // 2 - 1] + s[n // 2])) / 2.0)
        assert oracle.median(v) == expect
    assert oracle.median(np.array([], np.float32)) == 0.0
    assert oracle.median(np.array([7.0, 3.0], np.float32)) == 7.0        # n < 3 -> data[0] (utils.h:248-249)


def test_solve6_against_numpy(oracle):
    rng = np.random.RandomState(7)
    for _ in range(20):
        A = rng.randn(40, 6)
        H = (A.T @ A).astype(np.float32)
        G = rng.randn(6).astype(np.float32)
        ok, dp = oracle.solve6(H, G)
        assert ok
        ref = np.linalg.solve(H.astype(np.float64), G.astype(np.float64))
        assert np.abs(dp - ref).max() <= 2e-4 * np.abs(ref).max()
    # rank-deficient system: the float LDLT fails isApprox, the damped double retry (u = 1e-3 * max diag) takes over
    H = np.zeros((6, 6), np.float32); H[0, 0] = 1.0
    ok, dp = oracle.solve6(H, np.ones(6, np.float32))
    assert ok and abs(dp[0] - 1.0 / 1.001) < 1e-6 and np.allclose(dp[1:], 1000.0, rtol=1e-5)


def test_params_to_pose_is_conjugated_exponential(oracle):
    from scipy.linalg import expm
    p = np.array([0.01, -0.02, 0.015, 0.1, -0.05, 0.2], np.float32)
    Tn = np.eye(4, dtype=np.float32); Tn[:3, :3] *= 0.37; Tn[:3, 3] = [-0.5, 0.2, -3.0]
    xi = np.zeros((4, 4)); w, v = p[:3].astype(np.float64), p[3:].astype(np.float64)
    xi[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]; xi[:3, 3] = v
    expect = np.linalg.inv(Tn.astype(np.float64)) @ expm(xi) @ Tn.astype(np.float64)
    assert np.abs(oracle.params_to_pose(Tn, p) - expect).max() < 5e-6


def test_template_invariants(oracle):
    from bpvo_b200 import synth
    sc = synth.scene_small(96, 128)
    for desc in ["intensity", "bitplanes"]:
        p = make_params(desc, 3)
        f = oracle.Frame(sc.K, sc.baseline, sc.rows, sc.cols, p)
        img, d = sc.render(0)
        f.set_data(img, d); f.set_template()
        for l in range(3):
            n = f.num_points(l)
            assert n > 0 and n % 16 == 0                       # template_data.cc:85-89
            inds = f.point_inds(l)
            assert np.all(np.diff(inds) > 0)                   # scan order
            r, c = f.level_size(l)
            ys, xs = inds // c, inds % c
            assert ys.min() >= 3 and ys.max() < r - 4 and xs.min() >= 3 and xs.max() < c - 4
            pts = f.points(l)
            assert np.all(pts[:, 3] == 1.0) and np.all(pts[:, 2] > 0)
            Tn = f.normalization(l)
            q = (Tn @ pts.T).T[:, :3]
            assert np.abs(q.mean(axis=0)).max() < 1e-3 and abs(np.linalg.norm(q, axis=1).mean() - np.sqrt(3)) < 1e-3


def test_rcp_jacobians_stay_within_bound(oracle):
    """_mm_rcp_ps Jacobians (reference, rigid_body_warp.cc:47-58) vs exact division: <= 1.5 * 2^-12 * (a few ops)"""
    from bpvo_b200 import synth
    sc = synth.scene_small(96, 128)
    p = make_params("bitplanes", 2)
    img, d = sc.render(0)
    J = []
    for rcp in (0, 1):
        f = oracle.Frame(sc.K, sc.baseline, sc.rows, sc.cols, p, use_rcp=rcp)
        f.set_data(img, d); f.set_template()
        J.append(f.jacobians(0))
    scale = np.abs(J[0]).max(axis=(0, 1), keepdims=True)
    assert 0 < (np.abs(J[0] - J[1]) / scale).max() < 2e-3


@pytest.mark.parametrize("desc,levels,loss", [("intensity", 3, "huber"), ("bitplanes", 3, "tukey"), ("intensity", 2, "l2")])
def test_vo_recovers_synthetic_motion(oracle, desc, levels, loss):
    from bpvo_b200 import synth
    sc = synth.scene_small(120, 160)
    p = make_params(desc, levels, loss)
    vo = oracle.VisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p)
    evals = 0
    for k in range(5):
        img, d = sc.render(k)
        r = vo.add_frame(img, d)
        evals += r["numFunEvals"]
        if k == 0:
            assert r["isKeyFrame"] and r["keyFramingReason"] == 0x44 and np.array_equal(r["pose"], np.eye(4, dtype=np.float32))
        else:
            gt = sc.relative_pose(k - 1, k)
            # sanity only: on the 120x160 toy scene the motion is ~0.5 px per frame; census bit-planes resolve that
            # coarsely, intensity resolves it well
            tol_t, tol_r = (5e-3, 2e-3) if desc == "intensity" else (3e-2, 1e-2)
            assert np.abs(r["pose"][:3, 3] - gt[:3, 3]).max() < tol_t
            assert np.abs(r["pose"][:3, :3] - gt[:3, :3]).max() < tol_r
    assert evals > 0 and vo.trajectory().shape == (5, 4, 4)


def test_oracle_threads_agree(oracle):
    """the OpenMP stand-in for the TBB path must not change residuals / weights (only the reduction order)"""
    from bpvo_b200 import synth
    sc = synth.scene_small(96, 128)
    p = make_params("bitplanes", 2, "tukey")
    img0, d0 = sc.render(0); img1, d1 = sc.render(1)
    outs = []
    for nt in (1, 4):
        a = oracle.Frame(sc.K, sc.baseline, sc.rows, sc.cols, p, num_threads=nt)
        b = oracle.Frame(sc.K, sc.baseline, sc.rows, sc.cols, p, num_threads=nt)
        a.set_data(img0, d0); a.set_template(); b.set_data(img1, d1)
        e = oracle.Estimator(p, num_threads=nt)
        outs.append(e.linearize(a, b, 0, np.eye(4, dtype=np.float32)))
    assert np.array_equal(outs[0]["residuals"], outs[1]["residuals"]) and np.array_equal(outs[0]["weights"], outs[1]["weights"])
    assert np.abs(outs[0]["H"] - outs[1]["H"]).max() <= 1e-4 * np.abs(outs[0]["H"]).max()


def test_golden_stream_fixture(oracle):
    """pins the oracle itself across hosts/compilers: a committed fixture generated by tests/golden/make_stream_golden.py
    (exact-division mode so that it does not depend on the CPU vendor's rcpps)"""
    path = os.path.join(GOLD, "stream_small_bitplanes.npz")
    g = np.load(path)
    from bpvo_b200 import synth
    sc = synth.scene_small(96, 128)
    p = make_params("bitplanes", 3, "tukey")
    vo = oracle.VisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p, use_rcp=0)
    img0, _ = sc.render(0)
    assert np.array_equal(img0, g["image0"]), "the synthetic generator is not bit-reproducible on this host"
    for k in range(int(g["nframes"])):
        img, d = sc.render(k)
        r = vo.add_frame(img, d)
        assert r["isKeyFrame"] == bool(g["is_kf"][k])
        assert np.abs(r["pose"] - g["poses"][k]).max() < 1e-4
    f = vo.ref_frame()
    assert [f.num_points(l) for l in range(3)] == list(g["npoints"])


def test_general_gaussian_blur_and_gradient_descriptors_against_cv2(oracle):
    """third-party arithmetic of the two extra descriptors (bpvo/gradient_descriptor.cc): cv::GaussianBlur on CV_32F with the
    kernel size OpenCV derives from sigma (cv::Size()) or imsmooth's max(5, 2 round(sigma) + 1) -- the oracle's restatement
    against the real OpenCV (cv2 4.13) to a few ulp of the 0..255 range; then GradientDescriptor / DescriptorFields as numpy
    restatements of gradient_descriptor.cc:42-64, 78-116 on top of cv2's blur"""
    cv2 = pytest.importorskip("cv2")
    from conftest import make_params
    rng = np.random.RandomState(5)
    img = rng.randint(0, 256, size=(61, 83)).astype(np.uint8)
    f = img.astype(np.float32)
    for sigma in (0.75, 1.0, 1.2, 2.0):
        a = oracle.gaussian_blur_f32(f, 0, sigma)
        b = cv2.GaussianBlur(f, (0, 0), sigma, sigmaY=sigma, borderType=cv2.BORDER_REFLECT_101)
        assert np.abs(a - b).max() <= 1e-4, sigma
    for sigma, k in ((0.75, 5), (1.75, 5), (3.0, 7)):
        a = oracle.gaussian_blur_f32(f, k, sigma)
        b = cv2.GaussianBlur(f, (k, k), sigma, sigmaY=sigma, borderType=cv2.BORDER_REFLECT_101)
        assert np.abs(a - b).max() <= 1e-4, (sigma, k)

    def xgrad(I):
        g = np.empty_like(I)
        g[:, 0] = 0.5 * (I[:, 1] - I[:, 0]); g[:, 1:-1] = 0.5 * (I[:, 2:] - I[:, :-2]); g[:, -1] = 0.5 * (I[:, -1] - I[:, -2])
        return g

    def ygrad(I):
        return xgrad(I.T.copy()).T.copy()

    def smooth(I, s):
        k = max(5, 2 * int(round(s)) + 1)
        return cv2.GaussianBlur(I, (k, k), s, sigmaY=s, borderType=cv2.BORDER_REFLECT_101)

    for sig in (-1.0, 1.0):
        d = oracle.descriptor(make_params("gradient", 1, sigmaPriorToCensusTransform=sig), img)
        I = f if sig <= 0 else cv2.GaussianBlur(f, (0, 0), sig, sigmaY=sig, borderType=cv2.BORDER_REFLECT_101)
        assert d.shape == (3, 61, 83) and np.array_equal(d[0], f)
        assert np.abs(d[1] - xgrad(I)).max() <= 1e-4 and np.abs(d[2] - ygrad(I)).max() <= 1e-4
    d = oracle.descriptor(make_params("dfields", 1), img)
    I = smooth(f, 0.75)
    assert d.shape == (5, 61, 83) and np.array_equal(d[0], f)
    for c, g in ((1, xgrad(I)), (3, ygrad(I))):
        pos, neg = np.where(g >= 0, g, 0).astype(np.float32), np.where(g < 0, g, 0).astype(np.float32)
        assert np.abs(d[c] - smooth(pos, 1.75)).max() <= 2e-4 and np.abs(d[c + 1] - smooth(neg, 1.75)).max() <= 2e-4


# ---- upstream stereo: OpenCV's StereoBM as utils/stereo_algorithm.cc:67-111 configures and runs it ----
def _stereo_golden():
    g = np.load(os.path.join(GOLD, "stereo_bm.npz"))
    for i in range(int(g["n"])):
        nd, wsz, mind, cap, tex, uniq = (int(v) for v in g[f"params_{i}"])
        yield i, g[f"left_{i}"], g[f"right_{i}"], dict(numberOfDisparities=nd, SADWindowSize=wsz, minDisparity=mind, preFilterCap=cap,
                                                         textureThreshold=tex, uniquenessRatio=uniq), g[f"disp16_{i}"]


def test_stereo_bm_matches_cv2_golden(oracle):
    """cvFindStereoCorrespondenceBM (third-party; utils/stereo_algorithm.cc:107): the oracle's restatement is bit-exact against
    cv2 4.13 on the committed vectors (shifted noise, the synthetic rig, pure noise, a texture-less half, negative minDisparity)"""
    n = 0
    for i, left, right, kw, want in _stereo_golden():
        d16, df = oracle.stereo_bm(left, right, **kw)
        assert np.array_equal(d16, want), f"case {i}: {int((d16 != want).sum())} pixels differ"
        assert np.array_equal(df, want.astype(np.float32) / 16.0)        # convertTo(CV_32FC1, 1/16), stereo_algorithm.cc:110
        n += 1
    assert n >= 8


def test_stereo_bm_matches_cv2_live(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for trial in range(10):
        nd = int(rng.choice([16, 32, 64])); wsz = int(rng.choice([5, 9, 15, 21]))
        H = int(rng.integers(wsz + 8, 80)); W = int(rng.integers(nd + wsz + 8, nd + wsz + 100))
        mind = int(rng.choice([0, 0, -7, -nd])); tex = int(rng.choice([0, 10, 500])); uniq = int(rng.choice([0, 15, 30])); cap = int(rng.choice([5, 31, 63]))
        base = cv2.GaussianBlur(rng.integers(0, 256, (H, W + 2 * nd)).astype(np.uint8), (0, 0), 1.2)
        s = int(rng.integers(0, nd))
        left = base[:, nd:nd + W].copy(); right = base[:, nd + s:nd + s + W].copy()
        bm = cv2.StereoBM_create(nd, wsz)
        bm.setMinDisparity(mind); bm.setTextureThreshold(tex); bm.setUniquenessRatio(uniq); bm.setPreFilterCap(cap)
        d16, _ = oracle.stereo_bm(left, right, nd, wsz, mind, cap, tex, uniq)
        assert np.array_equal(d16, bm.compute(left, right)), (trial, H, W, nd, wsz, mind, tex, uniq, cap)


def test_stereo_bm_properties(oracle):
    """what the definition implies whatever the data: the border is filtered, a pure shift is recovered exactly, a
    texture-less image is filtered everywhere, unsupported parameters are refused"""
    from bpvo_b200.synth import scene_small
    rng = np.random.default_rng(9)
    nd, wsz, s = 32, 9, 11
    base = rng.integers(0, 256, (40, 200)).astype(np.uint8)
    left = base[:, 40:160].copy(); right = base[:, 40 + s:160 + s].copy()
    d16, df = oracle.stereo_bm(left, right, nd, wsz)
    inv = -16
    w2 = wsz // 2
    assert (d16[:w2] == inv).all() and (d16[-w2:] == inv).all() and (d16[:, :nd - 1 + w2] == inv).all() and (d16[:, -w2:] == inv).all()
    core = d16[w2:-w2, nd - 1 + w2:-w2]
    assert (np.abs(core.astype(int) - s * 16) <= 8).all()                # SAD 0 at the true shift; the sub-pixel step moves it by at most half a pixel
    flat = np.full((40, 120), 90, np.uint8)
    assert (oracle.stereo_bm(flat, flat, nd, wsz)[0] == inv).all()       # texture sum 0 < 10
    with pytest.raises(ValueError):
        oracle.stereo_bm(left, right, 30, wsz)                           # not a multiple of 16
    with pytest.raises(ValueError):
        oracle.stereo_bm(left, right, nd, 8)                             # even window
    with pytest.raises(ValueError):
        oracle.stereo_bm(left, right, nd, wsz, minDisparity=2)           # OpenCV writes out of bounds there: not restated
    # the synthetic rig: the right camera sits one baseline along x, so block matching recovers the rendered disparity
    sc = scene_small(rows=96, cols=160, seed=3); sc.baseline = 0.5
    L, D = sc.render(0); R = sc.render_right(0)
    d16, df = oracle.stereo_bm(L, R, 48, 9)
    ok = d16 >= 0
    assert ok.mean() > 0.4 and np.abs(df[ok] - D[ok]).mean() < 0.25
