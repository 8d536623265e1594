#!/bin/bash
# round 2, call E: whole GPU suite with diagnostics, sliced overflow scan, 8 streams
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TAG=${1:-r2e}
python -c "
from bpvo_b200 import _capi
L=_capi.lib(); print('device_count', L.bpvo_b200_device_count(), L.bpvo_b200_last_error())" > gpurun_out/${TAG}_diag.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q -rs > gpurun_out/${TAG}_pytest_gpu.log 2>&1
timeout 400 python bench.py --no-cpu-baseline --no-dense > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 300 python scripts/profile_kernels.py --workload 1080p_dense > gpurun_out/${TAG}_kernels_1080p_dense.json 2> gpurun_out/${TAG}_kernels_1080p_dense.err
timeout 300 python scripts/profile_kernels.py --workload kitti_dense > gpurun_out/${TAG}_kernels_dense.json 2> gpurun_out/${TAG}_kernels_dense.err
cat gpurun_out/${TAG}_diag.txt
grep -E "passed|failed|FAILED|^E  |SKIPPED" gpurun_out/${TAG}_pytest_gpu.log | tail -30
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("bench", d["value"], d["e2e"]["value"], d["gn_iters_per_frame"], [(t["streams_per_gpu"], round(t.get("value",0))) for t in d["throughput_mode"]])
for f in ("kernels_1080p_dense","kernels_dense"):
    d=json.load(open("gpurun_out/${TAG}_%s.json"%f))
    print(f, "hit", d["bracket_hit_rate"], [(L["level"], L["N"], L["us_per_gn_iter"], round(L["frac"],3)) for L in d["fused_levels"]])
PY
