#!/bin/bash
# round 2, call G: locate the fault of the TMA kernel (compute-sanitizer); everything else with BPVO_B200_NO_TMA=1
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TAG=${1:-r2g}
timeout 300 compute-sanitizer --tool memcheck python scripts/tma_debug.py > gpurun_out/${TAG}_tma_sanitizer.log 2>&1
export BPVO_B200_NO_TMA=1
timeout 2400 python -m pytest tests -m gpu -q -rs > gpurun_out/${TAG}_pytest_gpu.log 2>&1
for v in "" _noprefetch _small; do
  L=$PWD/bpvo_b200/libbpvo_b200$v.so
  BPVO_B200_LIB=$L timeout 300 python scripts/profile_kernels.py --workload 1080p_dense > gpurun_out/${TAG}_kernels_1080p_dense$v.json 2> gpurun_out/${TAG}_kernels_1080p_dense$v.err
  BPVO_B200_LIB=$L timeout 300 python scripts/profile_kernels.py --workload kitti_dense > gpurun_out/${TAG}_kernels_dense$v.json 2> gpurun_out/${TAG}_kernels_dense$v.err
  BPVO_B200_LIB=$L timeout 300 python bench.py --no-cpu-baseline --no-dense --no-throughput > gpurun_out/${TAG}_bench$v.json 2> gpurun_out/${TAG}_bench$v.err
done
head -60 gpurun_out/${TAG}_tma_sanitizer.log
grep -E "passed|failed|FAILED|^E  |SKIPPED" gpurun_out/${TAG}_pytest_gpu.log | tail -20
python - <<PY
import json
for v in ("","_noprefetch","_small"):
    d=json.load(open("gpurun_out/${TAG}_bench%s.json"%v)); print("bench"+v, round(d["value"],1), round(d["e2e"]["value"],1), d["gn_iters_per_frame"], round(1e3*d["ms_per_step"]/d["gn_iters_per_frame"],3))
    for f in ("kernels_1080p_dense","kernels_dense"):
        d=json.load(open("gpurun_out/${TAG}_%s%s.json"%(f,v)))
        print(f+v, "hit", round(d["bracket_hit_rate"],3), [(L["level"], L["N"], L["us_per_gn_iter"], round(L["frac"],3)) for L in d["fused_levels"]])
PY
