#!/bin/bash
# round 2, 8-GPU call: the driver's scaling command at N = 8 (replicas + sharded dense 1080p variant with parity fields)
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TAG=${1:-r2s}
N=${NGPU:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $N --steps 24 --warmup 4 > gpurun_out/${TAG}_bench$N.json 2> gpurun_out/${TAG}_bench$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench$N.json"))
print("N=$N value", d["value"], "e2e", d["e2e"]["value"])
s=d["sharded_1080p_dense_variant"]
print({k:s[k] for k in ("pose_rel_err_vs_single_gpu","parity_ok","speedup_vs_one_gpu","sharded_efficiency","translation_err_vs_ground_truth_m")})
print("single", s["single_gpu"]["us_per_gn_iter"], s["single_gpu"]["us_per_gn_iter_per_level"])
print("sharded", s["sharded"]["us_per_gn_iter"], s["sharded"]["us_per_gn_iter_per_level"], s["sharded"]["points_per_level_local"], s["sharded"]["poses_identical_across_ranks"])
PY
tail -3 gpurun_out/${TAG}_bench$N.err
