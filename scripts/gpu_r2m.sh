#!/bin/bash
# round 2, 2-GPU call: the two multi-GPU parity tests (NCCL point-sharded, peer-memory device loop) + bench at N=2 (replicas + sharded variant with parity fields)
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TAG=${1:-r2m}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_env.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -rs -k "two_gpus" > gpurun_out/${TAG}_pytest_2gpu.log 2>&1
tail -15 gpurun_out/${TAG}_pytest_2gpu.log | cut -c1-1500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 32 --warmup 4 > gpurun_out/${TAG}_bench2.json 2> gpurun_out/${TAG}_bench2.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench2.json"))
print("N=2 value", d["value"], "e2e", d["e2e"]["value"])
print(json.dumps(d["sharded_1080p_dense_variant"], indent=1)[:2500])
PY
tail -3 gpurun_out/${TAG}_bench2.err
