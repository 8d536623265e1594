#!/bin/bash
# 2-GPU call: sharded-mode check + replica bench at N=2 (as the driver launches it)
cd ${GRAFT_REPO_ROOT:-/root/repo}
TAG=${1:-r1m}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_env.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/shard_check.py > gpurun_out/${TAG}_shard.log 2>&1
tail -5 gpurun_out/${TAG}_shard.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 32 --warmup 4 > gpurun_out/${TAG}_bench2.json 2> gpurun_out/${TAG}_bench2.err
cat gpurun_out/${TAG}_bench2.json | cut -c1-600; tail -3 gpurun_out/${TAG}_bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 4 --warmup 1 > gpurun_out/${TAG}_bench2_ref.json 2> gpurun_out/${TAG}_bench2_ref.err
cut -c1-300 gpurun_out/${TAG}_bench2_ref.json
