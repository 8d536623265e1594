"""GPU parity of the PRODUCT kernel: the persistent on-device Gauss-Newton loop (k_estimate_pose) that
bpvo_b200_estimate_pose / bpvo_b200_vo_add_frame launch.

test_gpu_parity.py checks the vectors (r, valid, sigma, w, H, G) of the host-driven linearize kernels against the
oracle; here the SAME vectors are taken out of the persistent kernel -- its shared-memory template cache, its bracketed
exact median, its flag-in-data exchange, its device LDL^T -- and compared

  * with the fine seam (bpvo_b200_linearize):  r / valid / sigma / w bit-equal, H / G <= 1e-6 relative,
  * with the oracle: per-level OptimizerStatistics (status, numIterations, finalError) and the per-iteration trace
    {f_norm, |dp|, sigma} of PoseEstimatorBase::run (pose_estimator_base.h:324-407),
  * at BASELINE.json configs #1 (640x480 intensity, 1 level, kL2) and #4 (1920x1080 bit-planes, 5 levels, Tukey).
"""
import numpy as np
import pytest

from conftest import make_params, rel_err
from test_gpu_parity import _cmp_linearize, _pair

pytestmark = pytest.mark.gpu


def _poses(sc, n=9):
    """identity, then the ground-truth motion approached geometrically like a converging GN run: the median of |r| moves by
    percents at first (wide brackets, radix fallbacks), by 1e-4 at the end (narrow brackets)"""
    Tgt = np.array(sc.relative_pose(0, 1), dtype=np.float64)
    out = [np.eye(4, dtype=np.float32)]
    for k in range(1, n):
        T = Tgt.copy()
        T[:3, 3] += np.array([0.03, -0.02, 0.05]) * (0.4 ** k)
        out.append(T.astype(np.float32))
    return out


def _fine_seam_sequence(ctx, gref, gcur, level, poses):
    outs = []
    for k, T in enumerate(poses):
        outs.append(ctx.linearize(gref, gcur, level, T, k == 0))
    return outs, ctx.getResiduals(), ctx.getValidFlags(), ctx.getWeights()


def _check_device_vs_fine(ctx, gref, gcur, level, poses, grid_ctas=0, cache_bytes=-1, need_bracket=True):
    fine, r_f, v_f, w_f = _fine_seam_sequence(ctx, gref, gcur, level, poses)
    dev = ctx.debug_device_linearize(gref, gcur, level, poses, grid_ctas=grid_ctas, cache_bytes=cache_bytes)
    r_d, v_d, w_d = ctx.getResiduals(), ctx.getValidFlags(), ctx.getWeights()
    paths = [d["scale_path"] for d in dev]
    info = dict(level=level, N=gref.numPoints(level), grid=grid_ctas, cache=cache_bytes, scale_paths=paths)
    print("device-loop linearize vs fine seam:", info)
    for k, (f, d) in enumerate(zip(fine, dev)):
        assert d["n_valid"] == f["n_valid"], (k, info)
        assert np.float32(d["sigma"]).tobytes() == np.float32(f["sigma"]).tobytes(), (k, d["sigma"], f["sigma"], info)   # bit-equal
        assert rel_err(d["H"], f["H"]) <= 1e-6, (k, rel_err(d["H"], f["H"]), info)
        # G_a = sum w r J_a cancels towards the optimum; its rounding scales with sum |w r J_a| <= sqrt(H_aa) * f_norm
        gtol = 1e-6 * (np.sqrt(np.abs(np.diag(f["H"]))) * f["f_norm"] + np.abs(f["G"]))
        assert np.all(np.abs(d["G"] - f["G"]) <= gtol), (k, d["G"], f["G"], gtol, info)
        assert abs(d["f_norm"] - f["f_norm"]) <= 1e-6 * max(1.0, f["f_norm"]), (k, info)
    assert np.array_equal(v_d, v_f), info
    assert np.array_equal(r_d.view(np.uint32), r_f.view(np.uint32)), info           # residual vector bit-equal
    assert np.array_equal(w_d.view(np.uint32), w_f.view(np.uint32)), info           # weight vector bit-equal
    if need_bracket and ctx.params.lossFunction != 0x12 and any(p != 0 for p in paths[1:]):
        assert 3 in paths[1:], f"the bracketed median never hit: {info}"
    return dev


@pytest.mark.parametrize("kind,desc,levels,loss,kw", [
    ("small", "bitplanes", 3, "tukey", {}),
    ("small", "intensity", 3, "huber", {}),
    ("small", "intensity", 3, "l2", {}),
    ("odd", "bitplanes", 2, "huber", {}),
    ("vga", "intensity", 4, "huber", {}),
    ("kitti", "bitplanes", 4, "tukey", {}),                         # the headline workload (semi-dense)
    ("kitti", "bitplanes", 4, "tukey", {"nonMaxSuppRadius": -1}),   # dense: 11 cache slots per thread at level 0
    ("small", "gradient", 3, "huber", {}),                          # 3 channels, record stride 4
    ("vga", "dfields", 3, "tukey", {}),                             # 5 channels, record stride 8
    ("small", "bitplanes", 3, "tukey", {"_flags": 4}),              # BPVO_B200_FLAG_FAST_BLEND: the fp32-FMA blend instantiation
    ("kitti", "bitplanes", 4, "tukey", {"_flags": 4}),
])
def test_device_linearize_equals_fine_seam(kind, desc, levels, loss, kw, oracle):
    kw = dict(kw)
    flags = kw.pop("_flags", 0)
    p = make_params(desc, levels, loss, **kw)
    sc, ctx, gref, gcur, oref, ocur = _pair(kind, p, oracle, flags=flags)
    poses = _poses(sc)
    for l in range(levels - 1, -1, -1):
        # (dense selection: the jumpy test sequence gives brackets wider than the candidate buffers -> radix select throughout)
        _check_device_vs_fine(ctx, gref, gcur, l, poses, need_bracket=gref.numPoints(l) < 100000)


@pytest.mark.parametrize("grid,cache", [(1, -1), (3, 40 * 1024), (5, 12 * 1024), (148, 0), (16, 3 * 1024)])
def test_device_linearize_multi_slot_and_partial_cache(grid, cache, oracle):
    """few CTAs and little shared memory at a small size: several points per thread (K > 1 slots), fields that do not fit
    stay in global memory (tpl_cache_plan), the bracket's candidate regions overflow into the shared list"""
    p = make_params("bitplanes", 2, "tukey", nonMaxSuppRadius=-1)
    sc, ctx, gref, gcur, oref, ocur = _pair("small", p, oracle)
    poses = _poses(sc)
    for l in (1, 0):
        _check_device_vs_fine(ctx, gref, gcur, l, poses, grid_ctas=grid, cache_bytes=cache, need_bracket=False)


def test_device_linearize_1080p_dense_level0(oracle):
    """BASELINE config #4 on one GPU, level 0 of the dense selection: 1.83 M points, 49 slots per thread, only part of the
    fields fit the shared-memory cache; plus the oracle on the first evaluation"""
    p = make_params("bitplanes", 5, "tukey", nonMaxSuppRadius=-1)
    sc, ctx, gref, gcur, oref, ocur = _pair("1080p", p, oracle)
    assert gref.numPoints(0) > 1_500_000
    poses = _poses(sc, n=4)
    _check_device_vs_fine(ctx, gref, gcur, 0, poses, need_bracket=False)
    _check_device_vs_fine(ctx, gref, gcur, 4, poses, need_bracket=False)
    oest = oracle.Estimator(ctx.params)
    _cmp_linearize(ctx, gref, gcur, oref, ocur, oest, 0, poses[0], True, 3e-4)


def test_split_solve_equals_one_launch(oracle, monkeypatch):
    """dense 1080p: the finest level does not fit the shared-memory cache, so the solve is split by level -- the cached levels in
    the usual kernel, level 0 in the instantiation with two points per thread in flight, chained through the pose on the device.
    Same arithmetic per point, same summation order: pose, statistics and evaluation counts are IDENTICAL to the one-launch
    solve (BPVO_B200_NO_STREAM2), and so are the residuals / weights left behind by the last evaluation"""
    p = make_params("bitplanes", 5, "tukey", nonMaxSuppRadius=-1)
    sc, ctx, gref, gcur, oref, ocur = _pair("1080p", p, oracle)
    T0 = np.eye(4, dtype=np.float32)
    launches0 = ctx.counters()["launches"]
    Ta, sa, na = ctx.estimatePose(gref, gcur, T0)
    ra, wa, eva = np.array(ctx.getResiduals()), np.array(ctx.getWeights()), ctx.last_level_evals()
    launches_split = ctx.counters()["launches"] - launches0
    monkeypatch.setenv("BPVO_B200_NO_STREAM2", "1")
    Tb, sb, nb = ctx.estimatePose(gref, gcur, T0)
    rb, wb, evb = np.array(ctx.getResiduals()), np.array(ctx.getWeights()), ctx.last_level_evals()
    launches_one = ctx.counters()["launches"] - launches0 - launches_split
    assert launches_split == launches_one + 1                        # one more (cooperative) launch, nothing else
    assert np.array_equal(Ta, Tb) and na == nb and list(eva) == list(evb)
    for l in range(5):
        assert (sa[l].numIterations, sa[l].status, sa[l].finalError) == (sb[l].numIterations, sb[l].status, sb[l].finalError), l
    assert np.array_equal(ra, rb) and np.array_equal(wa, wb)


def _oracle_vs_gpu_solve(kind, p, oracle, pose_tol=1e-4):
    sc, ctx, gref, gcur, oref, ocur = _pair(kind, p, oracle, use_rcp=1)
    oest = oracle.Estimator(ctx.params)
    oest.set_trace(True)
    T0 = np.eye(4, dtype=np.float32)
    To, so, no = oest.estimate_pose(oref, ocur, T0)
    tr_o = oest.get_trace()
    ctx.set_trace(True)
    Tg, sg, ng = ctx.estimatePose(gref, gcur, T0)
    tr_g = ctx.get_trace()
    ev = ctx.last_level_evals()
    L = ctx.params.numPyramidLevels
    report = [dict(level=l, gpu=(sg[l].numIterations, hex(sg[l].status), sg[l].finalError),
                   oracle=(so[l]["numIterations"], hex(so[l]["status"]), so[l]["finalError"])) for l in range(L)]
    print("per-level OptimizerStatistics  (numIterations, status, finalError):", report, "evals gpu/oracle:", ng, no)
    assert rel_err(Tg, To) < pose_tol, (rel_err(Tg, To), report)
    for l in range(ctx.params.maxTestLevel, L):
        s = sg[l]
        assert s.status in (0x30, 0x31, 0x32, 0x33), report                      # a convergence verdict, never kSolverError
        assert 0 <= s.numIterations <= p.maxIterations, report
        if s.status == 0x33:                                                     # kMaxIterations: the cap was reached (Q2 bookkeeping)
            assert s.numIterations == p.maxIterations and ev[l] == s.numIterations + 2, (ev, report)
        else:                                                                    # converged on pass numIterations + 1
            assert ev[l] in (s.numIterations, s.numIterations + 1), (ev, report)
        # f at the optimum is stable to ~1e-4 relative whatever the iteration count
        assert abs(s.finalError - so[l]["finalError"]) <= 2e-3 * max(1.0, so[l]["finalError"]), report
    return sc, ctx, (Tg, sg, ng, tr_g), (To, so, no, tr_o), report


@pytest.mark.parametrize("kind,desc,levels,loss", [("small", "bitplanes", 3, "tukey"), ("vga", "intensity", 4, "huber"),
                                                    ("kitti", "bitplanes", 4, "tukey")])
def test_estimate_pose_statistics_and_trace(kind, desc, levels, loss, oracle):
    """per-level status / numIterations / finalError of the device loop against the oracle, and the first iterations of
    the coarsest level (both start from the same pose there) row by row: f_norm, sigma, |dp|"""
    p = make_params(desc, levels, loss)
    sc, ctx, (Tg, sg, ng, tr_g), (To, so, no, tr_o), report = _oracle_vs_gpu_solve(kind, p, oracle)
    top = levels - 1
    g = tr_g[tr_g[:, 0] == top]
    o = tr_o[tr_o[:, 0] == top]
    assert len(g) >= 3 and len(o) >= 3
    for k in range(min(6, len(g), len(o))):
        assert g[k, 1] == o[k, 1] == k + 1
        assert abs(g[k, 2] - o[k, 2]) <= 2e-4 * max(1.0, o[k, 2]), ("f_norm", k, g[k], o[k])
        assert abs(g[k, 5] - o[k, 5]) <= 2e-4 * max(1e-3, o[k, 5]), ("sigma", k, g[k], o[k])
        assert abs(g[k, 3] - o[k, 3]) <= 2e-2 * o[k, 3] + 1e-6, ("|dp|", k, g[k], o[k])


@pytest.mark.parametrize("kind,desc,levels,loss", [("vga", "intensity", 4, "huber"), ("kitti", "bitplanes", 4, "tukey")])
def test_iteration_counts_at_shipped_tolerances(kind, desc, levels, loss, oracle):
    """conf/kitti_bitplanes.cfg:4-5 / conf/perf_bitplanes.cfg:4-5 tolerances (1e-6 / 1e-4): the loop ends on a robust
    criterion, so the iteration counts of the device loop and the oracle agree closely (with the ctor defaults the loop ends
    when |dp| reaches its fp32 noise floor of ~1e-7 and the counts are rounding-level chaotic)"""
    p = make_params(desc, levels, loss, parameterTolerance=1e-6, functionTolerance=1e-4, maxIterations=100)
    sc, ctx, (Tg, sg, ng, tr_g), (To, so, no, tr_o), report = _oracle_vs_gpu_solve(kind, p, oracle)
    for l in range(levels):
        assert abs(sg[l].numIterations - so[l]["numIterations"]) <= max(3, 0.25 * so[l]["numIterations"]), report
    assert abs(ng - no) <= max(4, 0.15 * no), (ng, no, report)


@pytest.mark.parametrize("nms", [1, -1])
def test_config1_vga_intensity_1level_l2(nms, oracle):
    """BASELINE.json configs[0]: single 640x480 frame pair, intensity descriptor, 1 pyramid level, kL2 (apps/vo_perf.cc),
    semi-dense (N ~ 20k) and dense (N = 299 408)"""
    p = make_params("intensity", 1, "l2", nonMaxSuppRadius=nms)
    sc, ctx, gref, gcur, oref, ocur = _pair("vga", p, oracle, use_rcp=0)
    assert gref.numPoints(0) == oref.num_points(0)
    if nms < 0:
        assert gref.numPoints(0) > 290000      # dense: every pixel of the selection window with a valid disparity
    assert np.array_equal(gref.point_inds(0), oref.point_inds(0))
    assert np.array_equal(gref.points(0), oref.points(0))
    oest = oracle.Estimator(ctx.params)
    T = np.eye(4, dtype=np.float32)
    _cmp_linearize(ctx, gref, gcur, oref, ocur, oest, 0, T, True, 1e-4)
    _check_device_vs_fine(ctx, gref, gcur, 0, _poses(sc, 4))
    # whole solve vs the reference-faithful oracle
    sc, ctx, gref, gcur, oref, ocur = _pair("vga", p, oracle, use_rcp=1)
    oest = oracle.Estimator(ctx.params)
    To, so, no = oest.estimate_pose(oref, ocur, T)
    Tg, sg, ng = ctx.estimatePose(gref, gcur, T)
    assert rel_err(Tg, To) < 1e-4, (Tg, To)
    assert sg[0].status in (0x30, 0x31, 0x32, 0x33)
    gt = np.array(sc.relative_pose(0, 1))
    assert np.abs(Tg[:3, 3] - gt[:3, 3]).max() < 0.01


def test_config4_1080p_5levels_tukey_semidense(oracle):
    """BASELINE.json configs[3] on ONE GPU, default (NMS) selection: template parity at all five levels, one linearize per
    level and the whole coarse-to-fine solve against the oracle"""
    p = make_params("bitplanes", 5, "tukey")
    sc, ctx, gref, gcur, oref, ocur = _pair("1080p", p, oracle, use_rcp=0)
    oest = oracle.Estimator(ctx.params)
    T = np.eye(4, dtype=np.float32)
    for l in range(4, -1, -1):
        assert np.array_equal(gref.saliency(l), oref.saliency(l)), l
        assert np.array_equal(gref.point_inds(l), oref.point_inds(l)), l
        assert np.array_equal(gref.points(l), oref.points(l)), l
        assert np.array_equal(gref.pixels(l), oref.pixels(l)), l
        _cmp_linearize(ctx, gref, gcur, oref, ocur, oest, l, T, True, 2e-4)
    p = make_params("bitplanes", 5, "tukey")
    _oracle_vs_gpu_solve("1080p", p, oracle)


def test_config4_1080p_5levels_tukey_dense(oracle):
    """... and the dense selection (nonMaxSuppRadius = -1: 1.83 M / 434 k / 105 k / 25 k / 6 k points) that "justifies"
    sharding: selection parity at every level, linearize vs the oracle at levels 0 and 4, and the whole solve on the GPU
    against the ground-truth motion (the oracle needs minutes for this solve)"""
    p = make_params("bitplanes", 5, "tukey", nonMaxSuppRadius=-1)
    sc, ctx, gref, gcur, oref, ocur = _pair("1080p", p, oracle, use_rcp=0)
    oest = oracle.Estimator(ctx.params)
    T = np.array(sc.relative_pose(0, 1), dtype=np.float32)
    for l in range(4, -1, -1):
        assert gref.numPoints(l) == oref.num_points(l), l
        assert np.array_equal(gref.point_inds(l), oref.point_inds(l)), l
    for l in (4, 0):
        _cmp_linearize(ctx, gref, gcur, oref, ocur, oest, l, T, True, 3e-4)
    Tg, sg, ng = ctx.estimatePose(gref, gcur, np.eye(4, dtype=np.float32))
    gt = np.array(sc.relative_pose(0, 1))
    assert np.abs(Tg[:3, 3] - gt[:3, 3]).max() < 0.01 and np.abs(Tg[:3, :3] - gt[:3, :3]).max() < 1e-3
    for l in range(5):
        assert sg[l].status in (0x30, 0x31, 0x32, 0x33)
