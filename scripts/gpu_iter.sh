#!/bin/bash
# quick iteration call: GPU parity tests, semi-dense phase profile (product + fine-marks build), short bench
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TAG=${1:-it}
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python scripts/profile_kernels.py --workload kitti > gpurun_out/${TAG}_kernels_semidense.json 2> gpurun_out/${TAG}_kernels_semidense.err
BPVO_B200_LIB=$PWD/bpvo_b200/libbpvo_b200_fine.so timeout 300 python scripts/profile_kernels.py --workload kitti > gpurun_out/${TAG}_kernels_semidense_fine.json 2> gpurun_out/${TAG}_kernels_semidense_fine.err
timeout 300 python scripts/profile_kernels.py --workload kitti_dense > gpurun_out/${TAG}_kernels_dense.json 2> gpurun_out/${TAG}_kernels_dense.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -4 gpurun_out/${TAG}_pytest_gpu.log
python - <<P
import json
for f in ("${TAG}_kernels_semidense.json","${TAG}_kernels_semidense_fine.json","${TAG}_kernels_dense.json"):
    try:
        d=json.load(open("gpurun_out/"+f)); sp=d["solve_profile"]
        print(f, "us/eval", round(sp["us_per_eval"],2), "hit", round(d["bracket_hit_rate"],2), sp["phase_us_per_eval"], d.get("fine_us_per_eval"))
    except Exception as e: print(f, "ERR", e)
try:
    b=json.load(open("gpurun_out/${TAG}_bench.json")); print("bench", b["value"], b["e2e"]["value"], b["roofline"]["frac"], b.get("roofline_dense_variant",{}).get("us_per_gn_iter"))
except Exception as e: print("bench ERR", e)
P
tail -3 gpurun_out/${TAG}_bench.err gpurun_out/${TAG}_kernels_semidense.err
