cd ${GRAFT_REPO_ROOT:-/root/repo}
TAG=${1:-r1c}
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/${TAG}_pytest.log
timeout 300 python scripts/profile_kernels.py --workload kitti > gpurun_out/${TAG}_kernels_semidense.json 2> gpurun_out/${TAG}_kernels_semidense.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -12 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_kernels_semidense.json | python -c "import json,sys; d=json.load(sys.stdin); print(json.dumps(d['solve_profile']))"; tail -3 gpurun_out/${TAG}_kernels_semidense.err; cut -c1-700 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
