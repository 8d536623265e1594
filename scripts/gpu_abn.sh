#!/bin/bash
# same-box comparison of several builds of the library on the semi-dense phase profile: gpu_abn.sh "" g8 g16 ...
cd ${GRAFT_REPO_ROOT:-/root/repo}
for rep in 1 2 3; do
  for v in "$@"; do
    lib=$PWD/bpvo_b200/libbpvo_b200${v:+_$v}.so
    BPVO_B200_LIB=$lib timeout 300 python scripts/profile_kernels.py --workload kitti 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); sp=d['solve_profile']; print('${v:-product}', 'us/eval', round(sp['us_per_eval'],3), 'exchange', sp['phase_us_per_eval']['sync4'])"
  done
done
