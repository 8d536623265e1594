#!/bin/bash
# round 2, final bench lines at the last commit: default bench (all variants) + shipped tolerances
cd ${GRAFT_REPO_ROOT:-/root/repo}
tag=${1:-r2final}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 300 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --workload kitti_cfg --no-dense --no-stereo > gpurun_out/${tag}_bench_cfg.json 2> gpurun_out/${tag}_bench_cfg.err
python - <<PY
import json
for n in ("bench", "bench_cfg"):
    d=json.loads(open('gpurun_out/${tag}_%s.json' % n).read().strip().splitlines()[-1])
    print(n, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'pageable', d['e2e'].get('from_pageable_host_memory'), 'iters/frame', d['gn_iters_per_frame'], 'thr', [round(x['value']) for x in (d.get('throughput_mode') or [])])
    if d.get('upstream_stereo_variant'): print(' stereo', d['upstream_stereo_variant']['kernel_ms_per_pair'], d['upstream_stereo_variant']['pairs_to_poses']['frames_per_sec'], d['upstream_stereo_variant']['pairs_per_sec_e2e_pinned_host'])
PY
