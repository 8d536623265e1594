"""Drop-in test (Option A): the reference's OWN bpvo/vo.cc -- compiled unchanged from /root/reference by
oracle/Makefile `gpuseam` against the replacement headers integration/bpvo/{vo_frame,vo_pose_estimator}.h and linked with
integration/vo_b200_seam.cc + libbpvo_b200.so -- must walk a stream like the all-CPU reference (oracle/_ref, the same vo.cc
on its own CPU classes): identical key-frame decisions, poses within 1e-4 relative, the same point clouds.  The library is
prebuilt here and travels with the snapshot; the test is skipped where it is absent."""
import numpy as np
import pytest

from conftest import make_params, rel_err
from test_gpu_parity import _scene

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,desc,levels,loss,nframes", [("small", "bitplanes", 3, "tukey", 8), ("vga", "intensity", 4, "huber", 6),
                                                            ("kitti", "bitplanes", 4, "tukey", 8)])
def test_reference_vo_cc_on_the_gpu_seam(kind, desc, levels, loss, nframes, oracle):
    L = oracle.seam_lib()
    if L is None or oracle.ref_lib() is None:
        pytest.skip("oracle/_ref libraries not shipped")
    from bpvo_b200 import VisualOdometry
    p = make_params(desc, levels, loss)
    sc = _scene(kind)
    size = (sc.rows, sc.cols)
    v_seam = oracle.RefVisualOdometry(sc.K, sc.baseline, size, p, lib=L)       # reference vo.cc + GPU seam
    v_cpu = oracle.RefVisualOdometry(sc.K, sc.baseline, size, p)               # reference vo.cc + reference CPU classes
    v_shim = VisualOdometry(sc.K, sc.baseline, size, p)                        # this repo's restatement of vo.cc on the same C ABI
    n_kf = 0
    for k in range(nframes):
        img, d = sc.render(k)
        rs, rc, rg = v_seam.add_frame(img, d), v_cpu.add_frame(img, d), v_shim.addFrame(img, d)
        assert rs["isKeyFrame"] == rc["isKeyFrame"] and rs["keyFramingReason"] == rc["keyFramingReason"], f"frame {k}: key-frame decision"
        tol = 1e-4 if kind != "small" else 1e-3       # (the toy stream does not converge within maxIterations, see test_vo_stream)
        assert rel_err(rs["pose"], rc["pose"]) < tol, f"frame {k}\n{rs['pose']}\n{rc['pose']}"
        # the host shim of this repo restates vo.cc: same engine underneath -> same pose up to the 4x4 products' rounding
        assert rs["isKeyFrame"] == rg.isKeyFrame and rel_err(rs["pose"], rg.pose) < 2e-6, f"frame {k}"
        for l in range(levels):
            assert rs["stats"][l]["status"] in (0x30, 0x31, 0x32, 0x33, 0x34)
        assert v_seam.num_points_at_level() == v_cpu.num_points_at_level()
        assert rs["numPointCloud"] == rc["numPointCloud"]
        if rs["numPointCloud"]:
            n_kf += 1
            xs, ws, gs = v_seam.point_cloud(rs["numPointCloud"])
            xc, wc, gc = v_cpu.point_cloud(rc["numPointCloud"])
            assert np.array_equal(xs, xc) and np.array_equal(gs, gc)
            assert np.abs(ws - wc).max() < 2e-2                                 # weights at the (slightly different) converged pose
    assert rel_err(v_seam.trajectory(), v_cpu.trajectory()) < (2e-4 if kind != "small" else 5e-3)
    assert n_kf >= 1 or kind != "kitti"        # the KITTI stream key-frames every 3-4 frames (0.046 m per frame against 0.15 m)
