#!/bin/bash
# round 2, call B: device-loop parity tests after the projection fix, fp32-blend vs exact-blend A/B
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TAG=${1:-r2b}
timeout 1800 python -m pytest tests/test_gpu_device_loop.py -m gpu -q -s > gpurun_out/${TAG}_pytest_device_loop.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/${TAG}_pytest_parity.log 2>&1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_fast.json 2> gpurun_out/${TAG}_bench_fast.err
BPVO_B200_EXACT_BLEND=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_exact.json 2> gpurun_out/${TAG}_bench_exact.err
timeout 300 python scripts/profile_kernels.py --workload kitti > gpurun_out/${TAG}_kernels_semidense.json 2> gpurun_out/${TAG}_kernels_semidense.err
timeout 300 python scripts/profile_kernels.py --workload kitti_dense > gpurun_out/${TAG}_kernels_dense.json 2> gpurun_out/${TAG}_kernels_dense.err
timeout 300 python scripts/profile_kernels.py --workload 1080p_dense > gpurun_out/${TAG}_kernels_1080p_dense.json 2> gpurun_out/${TAG}_kernels_1080p_dense.err
grep -E "passed|failed|FAILED|Error" gpurun_out/${TAG}_pytest_device_loop.log | tail -30; tail -5 gpurun_out/${TAG}_pytest_parity.log
for f in fast exact; do python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_$f.json"))
print("$f", d["value"], d["e2e"]["value"], d["gn_iters_per_frame"], d["roofline"]["frac"], d.get("roofline_dense_variant"), d.get("roofline_hbm_bound_variant"))
PY
done
