#!/bin/bash
# round 2, call D: whole GPU suite (incl. the reference-vo.cc-on-the-GPU-seam test), throughput mode
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TAG=${1:-r2d}
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
timeout 400 python bench.py --no-cpu-baseline --no-dense > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 400 python bench.py --no-cpu-baseline --no-dense --workload kitti_cfg > gpurun_out/${TAG}_bench_cfg.json 2> gpurun_out/${TAG}_bench_cfg.err
grep -E "passed|failed|FAILED|^E  " gpurun_out/${TAG}_pytest_gpu.log | tail -30; tail -2 gpurun_out/${TAG}_smoke.log
for f in bench bench_cfg; do python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_$f.json"))
print("$f", d["value"], d["e2e"]["value"], d["gn_iters_per_frame"], d["throughput_mode"])
PY
done
tail -3 gpurun_out/${TAG}_bench.err
