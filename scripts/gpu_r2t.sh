#!/bin/bash
# round 2, call t: upstream stereo (N4) parity + bench variant, bracket-thread A/B, full GPU suite
cd ${GRAFT_REPO_ROOT:-/root/repo}
tag=${1:-r2t}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== stereo tests"; timeout 900 python -m pytest tests/test_gpu_stereo.py -x -q 2>&1 | tail -15
echo "== sanitizer (golden cases)"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_stereo.py -x -q -k "golden or errors" 2>&1 | tail -6
echo "== A/B bracket thread"; bash scripts/gpu_abn.sh "" br0
echo "== full gpu suite"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== bench"; timeout 900 python bench.py --steps 32 --warmup 4 --no-dense --no-throughput > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 400 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'iters/frame', d['gn_iters_per_frame'])
print(json.dumps(d['upstream_stereo_variant'], indent=1))
PY
