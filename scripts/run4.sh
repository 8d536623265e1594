cd ${GRAFT_REPO_ROOT:-/root/repo}
TAG=${1:-r1j}
timeout 300 python scripts/profile_kernels.py --workload kitti > gpurun_out/${TAG}_kernels_semidense.json 2> gpurun_out/${TAG}_kernels_semidense.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_kernels_semidense.json')); print(json.dumps(d['solve_profile']['phase_us_per_eval'])); print(d['solve_profile']['us_per_eval'], d.get('bracket_hit_rate')); print(json.dumps(d.get('fine_cycles_share')))"
tail -3 gpurun_out/${TAG}_kernels_semidense.err
