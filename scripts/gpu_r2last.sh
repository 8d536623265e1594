#!/bin/bash
# the GPU suite and the smoke test at the last commit of round 2, for the record
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -rs 2>&1 | tail -12 > gpurun_out/r2last_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke >> gpurun_out/r2last_pytest_gpu.log 2>&1
cat gpurun_out/r2last_pytest_gpu.log
