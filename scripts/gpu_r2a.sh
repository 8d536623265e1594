#!/bin/bash
# round 2, call A: new device-loop parity tests, iteration traces (device loop / host loop vs oracle), bench at default and shipped tolerances
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TAG=${1:-r2a}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_env.txt
(nproc; lscpu | grep "Model name") >> gpurun_out/${TAG}_env.txt
timeout 1500 python -m pytest tests/test_gpu_device_loop.py -m gpu -q -s 2>&1 | tail -150 > gpurun_out/${TAG}_pytest_device_loop.log
timeout 300 python scripts/iter_trace.py --frames 8 > gpurun_out/${TAG}_iter_trace.json 2> gpurun_out/${TAG}_iter_trace.err
timeout 300 python scripts/iter_trace.py --frames 4 --flags 1 > gpurun_out/${TAG}_iter_trace_hostloop.json 2> gpurun_out/${TAG}_iter_trace_hostloop.err
timeout 300 python scripts/iter_trace.py --frames 6 --workload kitti_cfg > gpurun_out/${TAG}_iter_trace_cfg.json 2> gpurun_out/${TAG}_iter_trace_cfg.err
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -30 > gpurun_out/${TAG}_pytest_parity.log
timeout 400 python bench.py --no-dense > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 400 python bench.py --workload kitti_cfg --no-dense > gpurun_out/${TAG}_bench_cfg.json 2> gpurun_out/${TAG}_bench_cfg.err
timeout 400 python bench.py --workload kitti_cfg --impl reference --steps 16 --warmup 1 > gpurun_out/${TAG}_bench_cfg_ref.json 2> gpurun_out/${TAG}_bench_cfg_ref.err
tail -40 gpurun_out/${TAG}_pytest_device_loop.log; tail -5 gpurun_out/${TAG}_pytest_parity.log; cat gpurun_out/${TAG}_bench.json | cut -c1-600; cat gpurun_out/${TAG}_bench_cfg.json | cut -c1-600; cat gpurun_out/${TAG}_bench_cfg_ref.json | cut -c1-400
