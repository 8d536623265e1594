#!/bin/bash
# round 2, call P: message-based median hand-over: parity suite + same-box A/B against the barrier-based hand-over
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TAG=${1:-r2p}
timeout 3000 python -m pytest tests -m gpu -q -rs > gpurun_out/${TAG}_pytest_gpu.log 2>&1
for v in "" _nomsg; do
  L=$PWD/bpvo_b200/libbpvo_b200$v.so
  BPVO_B200_LIB=$L timeout 300 python bench.py --no-cpu-baseline --no-dense --no-throughput > gpurun_out/${TAG}_bench$v.json 2> gpurun_out/${TAG}_bench$v.err
  BPVO_B200_LIB=$L timeout 300 python bench.py --no-cpu-baseline --no-dense --no-throughput --workload kitti_cfg > gpurun_out/${TAG}_bench_cfg$v.json 2> gpurun_out/${TAG}_bench_cfg$v.err
  BPVO_B200_LIB=$L timeout 300 python scripts/profile_kernels.py --workload kitti > gpurun_out/${TAG}_kernels_semidense$v.json 2> gpurun_out/${TAG}_kernels_semidense$v.err
done
BPVO_B200_LIB=$PWD/bpvo_b200/libbpvo_b200_fine.so timeout 300 python scripts/profile_kernels.py --workload kitti > gpurun_out/${TAG}_kernels_semidense_fine.json 2> gpurun_out/${TAG}_kernels_semidense_fine.err
grep -E "passed|failed|FAILED|^E  " gpurun_out/${TAG}_pytest_gpu.log | tail -20
python - <<PY
import json
for v in ("","_nomsg"):
    d=json.load(open("gpurun_out/${TAG}_bench%s.json"%v)); print("bench"+v, round(d["value"],1), round(d["e2e"]["value"],1), d["gn_iters_per_frame"], [(l["level"], round(l["us_per_gn_iter"],2)) for l in d["roofline"]["per_level"]])
    d=json.load(open("gpurun_out/${TAG}_bench_cfg%s.json"%v)); print("cfg"+v, round(d["value"],1), round(d["e2e"]["value"],1), d["gn_iters_per_frame"])
    d=json.load(open("gpurun_out/${TAG}_kernels_semidense%s.json"%v)); sp=d["solve_profile"]; print("  us/eval", round(sp["us_per_eval"],2), "hit", round(d["bracket_hit_rate"],3), {k:round(x,2) for k,x in sp["phase_us_per_eval"].items()})
d=json.load(open("gpurun_out/${TAG}_kernels_semidense_fine.json")); print(d.get("fine_us_per_eval"))
PY
