#!/bin/bash
# fine-marks profile only (semi-dense)
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TAG=${1:-fine}
BPVO_B200_LIB=$PWD/bpvo_b200/libbpvo_b200_fine.so timeout 300 python scripts/profile_kernels.py --workload kitti > gpurun_out/${TAG}_kernels_semidense_fine.json 2> gpurun_out/${TAG}_kernels_semidense_fine.err
python - <<P
import json
d=json.load(open("gpurun_out/${TAG}_kernels_semidense_fine.json")); sp=d["solve_profile"]
print("us/eval", round(sp["us_per_eval"],2), "hit", round(d["bracket_hit_rate"],2), sp["phase_us_per_eval"]); print(d.get("fine_us_per_eval"))
P
