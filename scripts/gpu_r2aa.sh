#!/bin/bash
# round 2, last call: the split-solve equivalence test, memcheck of the streaming instantiation at a small size, full GPU suite + smoke
cd ${GRAFT_REPO_ROOT:-/root/repo}
tag=${1:-r2aa}
mkdir -p gpurun_out
echo "== split solve == one launch"; timeout 900 python -m pytest tests/test_gpu_device_loop.py -x -q -k "split_solve or multi_slot" 2>&1 | tail -4
echo "== memcheck: streaming instantiation forced at a small size"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_device_loop.py -x -q -k "multi_slot and 148-0" 2>&1 | tail -5
echo "== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
