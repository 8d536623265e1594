#!/bin/bash
# one gpurun call: launch list + ncu --set full of the dominant kernel + per-kernel timings + bench
cd ${GRAFT_REPO_ROOT:-/root/repo}
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 300 python scripts/profile_kernels.py --workload kitti > gpurun_out/${TAG}_kernels_semidense.json 2> gpurun_out/${TAG}_kernels_semidense.err
timeout 300 python scripts/profile_kernels.py --workload kitti_dense > gpurun_out/${TAG}_kernels_dense.json 2> gpurun_out/${TAG}_kernels_dense.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-dense > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_estimate_pose -s 3 -c 1 -o gpurun_out/${TAG}_k_estimate_pose \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dense > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_residuals|k_reduce|bitplanes_kernel" -c 6 -o gpurun_out/${TAG}_fine_seam \
    python scripts/profile_kernels.py --workload kitti_dense --iters 1 > gpurun_out/${TAG}_ncu_fine.log 2>&1
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_ncu_full.log
