"""Generates the cv2 golden vectors that pin the oracle's restatement of the THIRD-PARTY block matcher the reference calls in
front of the path (cvFindStereoCorrespondenceBM at utils/stereo_algorithm.cc:107, configured at :67-85).
Run once in the build container (cv2 4.13.0); the GPU box never needs cv2.  Output: tests/golden/stereo_bm.npz (small).
Each case: left / right u8, the StereoBM parameters, and cv2's CV_16S disparity map."""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from bpvo_b200.synth import scene_small  # noqa: E402

rng = np.random.RandomState(0xB200)
out = {}
cases = []
# (rows, cols, ndisp, wsz, minDisparity, preFilterCap, textureThreshold, uniquenessRatio, kind)
specs = [(48, 96, 16, 5, 0, 31, 10, 15, "shift"), (41, 97, 32, 9, 0, 31, 10, 15, "shift"), (64, 128, 48, 15, 0, 31, 10, 15, "scene"),
         (37, 80, 16, 7, -3, 15, 0, 0, "shift"), (50, 100, 32, 11, -16, 63, 200, 40, "noise"), (33, 75, 16, 21, 0, 1, 10, 5, "flat"),
         (96, 160, 64, 9, 0, 31, 10, 15, "scene"), (45, 70, 16, 5, -16, 31, 10, 15, "shift")]
for i, (r, c, nd, wsz, mind, cap, tex, uniq, kind) in enumerate(specs):
    if kind == "scene":
        sc = scene_small(rows=r, cols=c, seed=11 + i)
        sc.baseline = 0.6 if nd >= 48 else 0.3           # disparities of 24 / 12 px at the plane's 3 m
        left, right = sc.render(0)[0], sc.render_right(0)
    else:
        base = cv2.GaussianBlur(rng.randint(0, 256, size=(r, c + 2 * nd)).astype(np.uint8), (0, 0), 1.5)
        s = int(rng.randint(0, nd))
        left = base[:, nd:nd + c].copy()
        right = base[:, nd + s:nd + s + c].copy()
        right = np.clip(right.astype(int) + rng.randint(-3, 4, size=right.shape), 0, 255).astype(np.uint8)
        if kind == "noise":
            right = rng.randint(0, 256, size=(r, c)).astype(np.uint8)
        if kind == "flat":
            left[:, : c // 2] = 77          # texture-less half: the texture threshold filters it
    bm = cv2.StereoBM_create(nd, wsz)
    bm.setMinDisparity(mind); bm.setPreFilterCap(cap); bm.setTextureThreshold(tex); bm.setUniquenessRatio(uniq)
    assert bm.getPreFilterType() == 1 and bm.getSpeckleWindowSize() == 0 and bm.getDisp12MaxDiff() == -1   # what the reference sets
    out[f"left_{i}"] = left
    out[f"right_{i}"] = right
    out[f"params_{i}"] = np.array([nd, wsz, mind, cap, tex, uniq], np.int32)
    out[f"disp16_{i}"] = bm.compute(left, right)
    print(i, (r, c, nd, wsz, mind, cap, tex, uniq, kind), "valid", int((out[f'disp16_{i}'] > (mind - 1) * 16).sum()))
out["n"] = np.int32(len(specs))
out["cv2_version"] = np.array(cv2.__version__)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "stereo_bm.npz"), **out)
print("wrote stereo_bm.npz")
