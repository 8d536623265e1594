#!/bin/bash
# 2-GPU call: peer-memory mode check (parity + timing)
cd ${GRAFT_REPO_ROOT:-/root/repo}
TAG=${1:-peer}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_env.txt; nvidia-smi topo -m >> gpurun_out/${TAG}_env.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29521 scripts/peer_check.py > gpurun_out/${TAG}_peer.log 2>&1
echo "rc=$?"; tail -25 gpurun_out/${TAG}_peer.log | cut -c1-2500
