#!/bin/bash
# round 2, call C: parity tests on the new tail (fixed-point exchange + warp-0 solve), A/B against the fp64-mailbox tail, per-level profile
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TAG=${1:-r2c}
timeout 1800 python -m pytest tests/test_gpu_device_loop.py -m gpu -q -s > gpurun_out/${TAG}_pytest_device_loop.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/${TAG}_pytest_parity.log 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-dense > gpurun_out/${TAG}_bench_new.json 2> gpurun_out/${TAG}_bench_new.err
BPVO_B200_LIB=$PWD/bpvo_b200/libbpvo_b200_oldtail.so timeout 300 python bench.py --no-cpu-baseline --no-dense > gpurun_out/${TAG}_bench_oldtail.json 2> gpurun_out/${TAG}_bench_oldtail.err
timeout 300 python bench.py --no-cpu-baseline --no-dense --workload kitti_cfg > gpurun_out/${TAG}_bench_cfg_new.json 2> gpurun_out/${TAG}_bench_cfg_new.err
timeout 300 python scripts/profile_kernels.py --workload kitti > gpurun_out/${TAG}_kernels_semidense.json 2> gpurun_out/${TAG}_kernels_semidense.err
BPVO_B200_LIB=$PWD/bpvo_b200/libbpvo_b200_fine.so timeout 300 python scripts/profile_kernels.py --workload kitti > gpurun_out/${TAG}_kernels_semidense_fine.json 2> gpurun_out/${TAG}_kernels_semidense_fine.err
timeout 300 python scripts/profile_kernels.py --workload kitti_dense > gpurun_out/${TAG}_kernels_dense.json 2> gpurun_out/${TAG}_kernels_dense.err
timeout 300 python scripts/profile_kernels.py --workload 1080p_dense > gpurun_out/${TAG}_kernels_1080p_dense.json 2> gpurun_out/${TAG}_kernels_1080p_dense.err
grep -E "passed|failed|FAILED|Error" gpurun_out/${TAG}_pytest_device_loop.log | tail -30; tail -5 gpurun_out/${TAG}_pytest_parity.log
for f in new oldtail cfg_new; do python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_$f.json"))
print("$f", d["value"], d["e2e"]["value"], d["gn_iters_per_frame"], 1e3*d["ms_per_step"]/d["gn_iters_per_frame"], d["roofline"]["frac"])
PY
done
