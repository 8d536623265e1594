#!/bin/bash
# one gpurun call for the record (round 2): GPU tests, smoke, bench (both arms, both tolerance regimes), per-level profiles of the fused
# kernel, ncu launch list, ncu --set full of the dominant kernel (headline workload) and of the fused kernel on the dense 1080p workload
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TAG=${1:-r2}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_env.txt
(nproc; lscpu | grep "Model name") >> gpurun_out/${TAG}_env.txt
timeout 2400 python -m pytest tests -m gpu -q -rs 2>&1 | tail -40 > gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 16 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
timeout 600 python bench.py --workload kitti_cfg --no-dense --no-stereo > gpurun_out/${TAG}_bench_cfg.json 2> gpurun_out/${TAG}_bench_cfg.err
timeout 600 python bench.py --workload kitti_cfg --impl reference --steps 16 --warmup 1 > gpurun_out/${TAG}_bench_cfg_ref.json 2> gpurun_out/${TAG}_bench_cfg_ref.err
timeout 300 python scripts/profile_kernels.py --workload kitti > gpurun_out/${TAG}_kernels_semidense.json 2> gpurun_out/${TAG}_kernels_semidense.err
BPVO_B200_LIB=$PWD/bpvo_b200/libbpvo_b200_fine.so timeout 300 python scripts/profile_kernels.py --workload kitti > gpurun_out/${TAG}_kernels_semidense_fine.json 2> gpurun_out/${TAG}_kernels_semidense_fine.err
timeout 300 python scripts/profile_kernels.py --workload kitti_dense > gpurun_out/${TAG}_kernels_dense.json 2> gpurun_out/${TAG}_kernels_dense.err
timeout 300 python scripts/profile_kernels.py --workload 1080p_dense > gpurun_out/${TAG}_kernels_1080p_dense.json 2> gpurun_out/${TAG}_kernels_1080p_dense.err
timeout 300 python scripts/profile_stereo.py kitti > gpurun_out/${TAG}_stereo_kitti.json 2> gpurun_out/${TAG}_stereo.err
timeout 300 python scripts/profile_stereo.py 1080p > gpurun_out/${TAG}_stereo_1080p.json 2>> gpurun_out/${TAG}_stereo.err
BPVO_B200_NO_STREAM2=1 timeout 300 python scripts/profile_kernels.py --workload 1080p_dense > gpurun_out/${TAG}_kernels_1080p_dense_one_launch.json 2> gpurun_out/${TAG}_kernels_1080p_dense_one_launch.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-dense --no-throughput --no-stereo > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_estimate_pose -s 3 -c 1 -o gpurun_out/${TAG}_k_estimate_pose \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dense --no-throughput --no-stereo > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_k_estimate_pose.ncu-rep --page details > gpurun_out/${TAG}_k_estimate_pose_details.txt 2>&1
ncu -i gpurun_out/${TAG}_k_estimate_pose.ncu-rep --page raw --csv > gpurun_out/${TAG}_k_estimate_pose_raw.csv 2>&1
# the fused kernel where it is HBM bound: one solve of the dense 1080p workload (all five levels in the one launch)
timeout 900 ncu --set full --clock-control none -k regex:k_estimate_pose -s 1 -c 1 -o gpurun_out/${TAG}_k_estimate_pose_1080p_dense \
    python scripts/profile_kernels.py --workload 1080p_dense --iters 1 > gpurun_out/${TAG}_ncu_full_1080p.log 2>&1
ncu -i gpurun_out/${TAG}_k_estimate_pose_1080p_dense.ncu-rep --page raw --csv > gpurun_out/${TAG}_k_estimate_pose_1080p_dense_raw.csv 2>&1
tail -5 gpurun_out/${TAG}_pytest_gpu.log; tail -2 gpurun_out/${TAG}_smoke.log; cut -c1-700 gpurun_out/${TAG}_bench.json; tail -n 3 gpurun_out/${TAG}_bench.err; cut -c1-400 gpurun_out/${TAG}_bench_ref.json; tail -n 3 gpurun_out/${TAG}_ncu_full.log
