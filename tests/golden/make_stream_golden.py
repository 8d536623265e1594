"""Generates tests/golden/stream_small_bitplanes.npz: outputs of the oracle (exact-division mode) on a
deterministic synthetic stream.  Pins the oracle + generator across hosts; run in the build container."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bpvo_b200 import synth  # noqa: E402
from conftest import make_params  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

sc = synth.scene_small(96, 128)
p = make_params("bitplanes", 3, "tukey")
vo = po.VisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p, use_rcp=0)
poses, kf = [], []
N = 6
for k in range(N):
    img, d = sc.render(k)
    r = vo.add_frame(img, d)
    poses.append(r["pose"]); kf.append(r["isKeyFrame"])
f = vo.ref_frame()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "stream_small_bitplanes.npz"),
                    image0=sc.render(0)[0], poses=np.stack(poses), is_kf=np.array(kf), nframes=N,
                    npoints=np.array([f.num_points(l) for l in range(3)]))
print("ok", [f.num_points(l) for l in range(3)], kf)
