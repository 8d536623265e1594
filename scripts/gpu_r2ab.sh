#!/bin/bash
# round 2: stereo pipeline with the staged left image reused; memcheck of the streaming instantiation (grid 1: 40 slots per thread)
cd ${GRAFT_REPO_ROOT:-/root/repo}
tag=${1:-r2ab}
mkdir -p gpurun_out
echo "== stereo tests"; timeout 900 python -m pytest tests/test_gpu_stereo.py -x -q 2>&1 | tail -3
echo "== memcheck streaming instantiation"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_device_loop.py -x -q -k "multi_slot and 1--1" 2>&1 | tail -4
echo "== bench (short)"; timeout 900 python bench.py --steps 32 --warmup 4 --no-dense --no-throughput --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
s=d['upstream_stereo_variant']
print('value', d['value'], 'e2e', d['e2e']['value'], 'pairs->poses', s['pairs_to_poses'], 'e2e pairs/s', s['pairs_per_sec_e2e_pinned_host'], 'resident', s['pairs_per_sec_resident'])
PY
