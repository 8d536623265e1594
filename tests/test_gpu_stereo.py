"""GPU parity of the upstream stereo step: bpvo_b200_stereo_* (bpvo_b200/csrc/stereo.cu, through the C ABI)
against the CPU oracle's restatement of OpenCV's StereoBM (oracle/stereo_oracle.cc, itself bit-exact against cv2 4.13) and against
the committed cv2 golden vectors.  Integer work: every comparison is BIT-EXACT."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, make_params

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


def _stereo(shape, **kw):
    from bpvo_b200.stereo import StereoAlgorithm
    return StereoAlgorithm(shape, **kw)


def _pair(rng, H, W, nd, kind="shift"):
    base = rng.integers(0, 256, (H + 8, W + 2 * nd + 8)).astype(np.float64)
    # cheap smoothing (no cv2 on the GPU box needed): 3x3 box twice
    for _ in range(2):
        base = (base[:-2, :-2] + base[:-2, 1:-1] + base[:-2, 2:] + base[1:-1, :-2] + base[1:-1, 1:-1] + base[1:-1, 2:] + base[2:, :-2] + base[2:, 1:-1] + base[2:, 2:]) / 9.0
    base = np.clip(np.rint((base - 128) * 2.5 + 128), 0, 255).astype(np.uint8)
    s = int(rng.integers(0, nd))
    left = base[:H, nd:nd + W].copy(); right = base[:H, nd + s:nd + s + W].copy()
    right = np.clip(right.astype(int) + rng.integers(-4, 5, right.shape), 0, 255).astype(np.uint8)
    if kind == "noise":
        right = rng.integers(0, 256, (H, W)).astype(np.uint8)
    if kind == "flat":
        left[:, W // 3: 2 * W // 3] = 120
    return left, right


def test_matches_cv2_golden_vectors():
    g = np.load(os.path.join(GOLD, "stereo_bm.npz"))
    for i in range(int(g["n"])):
        nd, wsz, mind, cap, tex, uniq = (int(v) for v in g[f"params_{i}"])
        left, right, want = g[f"left_{i}"], g[f"right_{i}"], g[f"disp16_{i}"]
        st = _stereo(left.shape, numberOfDisparities=nd, SADWindowSize=wsz, minDisparity=mind, preFilterCap=cap, textureThreshold=tex, uniquenessRatio=uniq)
        dmap, d16 = st.run(left, right, want_fixed_point=True)
        assert np.array_equal(d16, want), f"case {i}: {int((d16 != want).sum())} pixels differ from cv2"
        assert np.array_equal(dmap, want.astype(np.float32) / 16.0)
        assert st.filteredValue() == float(mind - 1) and st.getInvalidValue() == float(mind - 1) / 16.0      # (the reference's getInvalidValue quirk)
        st.close()


@pytest.mark.parametrize("seed", range(6))
def test_random_configurations_match_the_oracle(oracle, seed):
    """odd sizes, windows 5..31, 16..256 disparities, negative minDisparity (clamped right-image columns), every threshold"""
    rng = np.random.default_rng(100 + seed)
    for trial in range(5):
        nd = int(rng.choice([16, 32, 64, 128, 256])); wsz = int(rng.choice([5, 7, 9, 11, 15, 21, 31]))
        H = int(rng.integers(wsz + 3, 150)); W = int(rng.integers(nd + wsz + 5, nd + wsz + 200))
        mind = int(rng.choice([0, 0, 0, -5, -nd // 2, -nd, -nd - 9])); tex = int(rng.choice([0, 10, 300, 3000])); uniq = int(rng.choice([0, 5, 15, 50]))
        cap = int(rng.choice([1, 7, 31, 63]))
        left, right = _pair(rng, H, W, nd, kind=["shift", "shift", "noise", "flat"][trial % 4])
        kw = dict(numberOfDisparities=nd, SADWindowSize=wsz, minDisparity=mind, preFilterCap=cap, textureThreshold=tex, uniquenessRatio=uniq)
        want16, wantf = oracle.stereo_bm(left, right, **kw)
        st = _stereo((H, W), **kw)
        dmap, d16 = st.run(left, right, want_fixed_point=True)
        pl, pr = st.prefiltered()
        assert np.array_equal(pl, oracle.stereo_prefilter_xsobel(left, cap)) and np.array_equal(pr, oracle.stereo_prefilter_xsobel(right, cap)), kw
        assert np.array_equal(d16, want16), (kw, H, W, int((d16 != want16).sum()))
        assert np.array_equal(dmap, wantf)
        st.close()


def test_kitti_size_with_the_reference_config(oracle):
    """conf/kitti.cfg:7-13 (BlockMatching, SADWindowSize 9, 128 disparities) on the synthetic KITTI-sized rig: bit-exact against the
    oracle, close to the rendered ground-truth disparity, repeatable, and the same through device-resident buffers"""
    from bpvo_b200.synth import scene_kitti
    sc = scene_kitti()
    L, D = sc.render(0); R = sc.render_right(0)
    cfg = {"StereoAlgorithm": "BlockMatching", "SADWindowSize": 9, "minDisparity": 0, "numberOfDisparities": 128, "trySmallerWindows": 1}
    st = _stereo(L.shape, config=cfg)
    dmap, d16 = st.run(L, R, want_fixed_point=True)
    want16, wantf = oracle.stereo_bm(L, R, 128, 9)
    assert np.array_equal(d16, want16) and np.array_equal(dmap, wantf)
    ok = d16 >= 0
    assert ok.mean() > 0.8 and np.abs(dmap[ok] - D[ok]).mean() < 0.1
    dmap2 = st.run(L, R)
    assert np.array_equal(dmap2, dmap)
    assert st.launches() == 6 and 0.0 < st.last_kernel_ms() < 50.0
    torch = pytest.importorskip("torch")
    dl, dr = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    out = torch.empty(L.shape, dtype=torch.float32, device="cuda"); out16 = torch.empty(L.shape, dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    st.run_raw(dl.data_ptr(), dr.data_ptr(), out.data_ptr(), out16.data_ptr())
    assert np.array_equal(out.cpu().numpy(), wantf) and np.array_equal(out16.cpu().numpy(), want16)
    st.close()


def test_1080p_properties():
    """full size of configs[3]: size-independent properties instead of the oracle -- a pure shift is recovered inside the valid
    region, the border is filtered, a texture-less pair is filtered everywhere"""
    rng = np.random.default_rng(3)
    H, W, nd, wsz, s = 1080, 1920, 128, 15, 37
    base = rng.integers(0, 256, (H, W + nd)).astype(np.uint8)
    left = base[:, :W].copy(); right = base[:, s:s + W].copy()
    st = _stereo((H, W), numberOfDisparities=nd, SADWindowSize=wsz)
    dmap, d16 = st.run(left, right, want_fixed_point=True)
    w2 = wsz // 2
    assert (d16[:w2] == -16).all() and (d16[-w2:] == -16).all() and (d16[:, :nd - 1 + w2] == -16).all() and (d16[:, -w2:] == -16).all()
    core = d16[w2:-w2, nd - 1 + w2:W - w2]
    assert (np.abs(core.astype(int) - s * 16) <= 8).all()
    flat = np.full((H, W), 77, np.uint8)
    assert (st.run(flat, flat) == -1.0).all()
    st.close()


def test_errors_mirror_opencv_and_the_reference():
    from bpvo_b200 import Error
    with pytest.raises(Error, match="numberOfDisparities"):
        _stereo((64, 128))
    with pytest.raises(Error, match="divisble by 16"):
        _stereo((64, 128), numberOfDisparities=24)
    with pytest.raises(Error, match="SADWindowSize must be odd"):
        _stereo((64, 128), numberOfDisparities=16, SADWindowSize=8)
    with pytest.raises(Error, match="SADWindowSize must be odd"):
        _stereo((20, 128), numberOfDisparities=16, SADWindowSize=21)          # not smaller than the image
    with pytest.raises(Error, match="preFilterCap"):
        _stereo((64, 128), numberOfDisparities=16, preFilterCap=64)
    for bad in (dict(minDisparity=3), dict(speckleWindowSize=50, speckleRange=2), dict(disp12MaxDiff=1), dict(preFilterType=0), dict(SADWindowSize=33)):
        with pytest.raises(Error, match=r"\[-5\]"):
            _stereo((64, 128), numberOfDisparities=16, **bad)
    with pytest.raises(Error, match="Unknown stereo algorithm"):
        _stereo((64, 128), numberOfDisparities=16, StereoAlgorithm="Magic")
    st = _stereo((64, 128), numberOfDisparities=16)
    with pytest.raises(Error, match="nullptr image"):
        st.run_raw(0, 0, 0)
    with pytest.raises(ValueError):
        st.run(np.zeros((10, 10), np.uint8), np.zeros((10, 10), np.uint8))
    # a window wider than the disparity-free part of the image: everything filtered, like OpenCV
    st2 = _stereo((40, 40), numberOfDisparities=48, SADWindowSize=5)
    assert (st2.run(np.zeros((40, 40), np.uint8), np.zeros((40, 40), np.uint8)) == -1.0).all()


def test_stereo_pairs_into_visual_odometry(oracle):
    """image pairs in, poses out (utils/dataset.cc:133 -> VisualOdometry::addFrame): addStereoFrame keeps the disparity map on the
    device and gives the pose addFrame gives for the same (bit-identical) disparity map passed from the host"""
    from bpvo_b200 import VisualOdometry
    from bpvo_b200.synth import scene_vga
    sc = scene_vga(); sc.baseline = 0.25                         # ~38 px of disparity at the plane's 4 m
    p = make_params("bitplanes", levels=4, loss="tukey", minValidDisparity=1.0)
    vo_pairs = VisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p)
    vo_host = VisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p)
    st = _stereo((sc.rows, sc.cols), numberOfDisparities=64, SADWindowSize=9)
    for k in range(4):
        L = sc.render(k)[0]; R = sc.render_right(k)
        r1 = vo_pairs.addStereoFrame(st, L, R, want_cloud=False)
        want16, wantf = oracle.stereo_bm(L, R, 64, 9)
        r2 = vo_host.addFrame(L, wantf)
        assert np.array_equal(r1.pose, r2.pose) and r1.isKeyFrame == r2.isKeyFrame and r1.numFunEvals == r2.numFunEvals
        if k > 0:
            gt = sc.relative_pose(k - 1, k)
            assert np.abs(r1.pose[:3, 3] - gt[:3, 3]).max() < 5e-3, (k, r1.pose[:3, 3], gt[:3, 3])
    st.close()


def test_cpp_stereo_mirror(tmp_path, oracle):
    """examples/stereo_vo_stream.cpp: bpvo_b200::StereoAlgorithm + bpvo_b200::VisualOdometry (the C++ mirrors of
    utils/stereo_algorithm.h and bpvo/vo.h) walk a stream of image pairs to the disparity maps of the oracle and the poses
    of the Python binding of the same C ABI"""
    import struct, subprocess
    from bpvo_b200 import VisualOdometry
    from bpvo_b200.synth import scene_small
    sc = scene_small(rows=120, cols=200, seed=5); sc.baseline = 0.4
    nframes, nd, wsz, levels = 4, 32, 9, 3
    pairs = [(sc.render(k)[0], sc.render_right(k)) for k in range(nframes)]
    binf = tmp_path / "pairs.bin"
    with open(binf, "wb") as f:
        f.write(struct.pack("<6i", sc.rows, sc.cols, nframes, nd, wsz, levels))
        f.write(np.asarray(sc.K, np.float32).T.tobytes()); f.write(struct.pack("<f", sc.baseline))
        for L, R in pairs:
            f.write(L.tobytes()); f.write(R.tobytes())
    exe = tmp_path / "stereo_vo_stream"
    libdir = os.path.join(ROOT, "bpvo_b200")
    subprocess.run(["/usr/bin/g++", "-std=c++14", "-O2", os.path.join(ROOT, "examples", "stereo_vo_stream.cpp"),
                    "-I" + os.path.join(libdir, "csrc", "host"), "-L" + libdir, "-lbpvo_b200", "-Wl,-rpath," + libdir, "-o", str(exe)], check=True)
    out = subprocess.run([str(exe), str(binf)], check=True, capture_output=True, text=True, timeout=300).stdout.splitlines()
    assert len(out) == nframes
    from bpvo_b200 import AlgorithmParameters, DescriptorType
    q = AlgorithmParameters(); q.descriptor = DescriptorType.kBitPlanes; q.numPyramidLevels = levels; q.minValidDisparity = 1.0
    vg = VisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), q)
    for k, (L, R) in enumerate(pairs):
        d16, df = oracle.stereo_bm(L, R, nd, wsz)
        tok = out[k].split()
        assert int(tok[3]) == int((d16 == -16).sum()) and float(tok[2]) == float(df[d16 != -16].astype(np.float64).sum())
        r = vg.addFrame(L, df)
        pose = np.array([float.fromhex(t) for t in tok[4:20]], np.float32).reshape(4, 4).T
        assert int(tok[0]) == int(r.isKeyFrame) and int(tok[1]) == r.numFunEvals and np.array_equal(pose, r.pose), f"frame {k}"
