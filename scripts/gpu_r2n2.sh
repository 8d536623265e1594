#!/bin/bash
# 2-GPU parity tests only (NCCL point-sharded host loop, peer-memory device loop)
cd ${GRAFT_REPO_ROOT:-/root/repo}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -rs -k "two_gpus" 2>&1 | tail -5
