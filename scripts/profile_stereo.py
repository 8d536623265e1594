"""Upstream stereo on the KITTI-sized synthetic pair with conf/kitti.cfg's settings: a few runs for ncu
(`ncu --set full -k regex:k_bm -c 3 python scripts/profile_stereo.py`) and, run plainly, the device time of the three kernels."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bpvo_b200.stereo import StereoAlgorithm  # noqa: E402
from bpvo_b200.synth import scene_kitti, scene_1080p  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "kitti"
sc = scene_kitti() if which == "kitti" else scene_1080p()
if which != "kitti":
    sc.baseline = 0.5
nd, wsz = (128, 9) if which == "kitti" else (128, 15)
L, R = sc.render(0)[0], sc.render_right(0)
st = StereoAlgorithm(L.shape, numberOfDisparities=nd, SADWindowSize=wsz)
ms = []
for _ in range(int(os.environ.get("REPS", "12"))):
    d = st.run(L, R)
    ms.append(st.last_kernel_ms())
cells = (sc.rows - wsz + 1) * (sc.cols - nd + 1 - wsz + 1) * nd
print(json.dumps({"workload": which, "rows": sc.rows, "cols": sc.cols, "numberOfDisparities": nd, "SADWindowSize": wsz,
                  "kernel_ms_median": float(np.median(ms[2:])), "kernel_ms_min": float(np.min(ms[2:])),
                  "giga_cells_per_sec": cells / (float(np.median(ms[2:])) * 1e-3) / 1e9, "valid_fraction": float((d >= 0).mean())}))
