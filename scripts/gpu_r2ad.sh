#!/bin/bash
# round 2: disparity upload on a copy stream beside the descriptor kernels and the solve: full GPU suite + bench lines (both tolerance regimes)
cd ${GRAFT_REPO_ROOT:-/root/repo}
tag=${1:-r2ad}
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
for w in kitti kitti_cfg; do
  timeout 600 python bench.py --workload $w --steps 64 --warmup 4 --no-dense --no-throughput --no-cpu-baseline --no-stereo > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench_$w.json').read().strip().splitlines()[-1])
print('$w', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ratio', round(d['e2e']['value']/d['value'],3))
PY
done
