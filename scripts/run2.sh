cd ${GRAFT_REPO_ROOT:-/root/repo}
timeout 900 python -m pytest tests -m gpu -q --tb=short -k "linearize or estimate_pose or vo_stream or full_size" -s 2>&1 | grep -E "linearize parity|passed|failed|FAILED|Error|assert" | head -150 > gpurun_out/r1b_pytest.log
timeout 300 python scripts/profile_kernels.py --workload kitti > gpurun_out/r1b_kernels_semidense.json 2> gpurun_out/r1b_kernels_semidense.err
tail -30 gpurun_out/r1b_pytest.log; cat gpurun_out/r1b_kernels_semidense.json; tail -3 gpurun_out/r1b_kernels_semidense.err
