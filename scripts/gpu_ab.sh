#!/bin/bash
# same-box A/B of two builds of the library: alternating runs of the semi-dense / dense phase profiles
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
A=${1:-prev}; B=${2:-}
for rep in 1 2 3; do
  for v in "$A" "$B"; do
    lib=$PWD/bpvo_b200/libbpvo_b200${v:+_$v}.so
    for w in kitti kitti_dense; do
      BPVO_B200_LIB=$lib timeout 300 python scripts/profile_kernels.py --workload $w 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); sp=d['solve_profile']; print('${v:-new}', '$w', 'us/eval', round(sp['us_per_eval'],2), 'hit', round(d['bracket_hit_rate'],2))"
    done
  done
done
