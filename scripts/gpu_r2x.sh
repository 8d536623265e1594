#!/bin/bash
# round 2, call x: solve split by level (cached levels in the usual kernel, the streaming level in the two-in-flight instantiation)
cd ${GRAFT_REPO_ROOT:-/root/repo}
tag=${1:-r2x}
mkdir -p gpurun_out
echo "== parity (device loop + streams)"; timeout 1500 python -m pytest tests/test_gpu_device_loop.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4
echo "== 1080p dense fused levels: split on / off"
for rep in 1 2; do
  timeout 600 python scripts/profile_kernels.py --workload 1080p_dense > gpurun_out/${tag}_kernels_1080p_dense.json 2> gpurun_out/${tag}_k.err
  BPVO_B200_NO_STREAM2=1 timeout 600 python scripts/profile_kernels.py --workload 1080p_dense > gpurun_out/${tag}_kernels_1080p_dense_nostream2.json 2>> gpurun_out/${tag}_k.err
  python - <<PY
import json
for name in ("", "_nostream2"):
    d = json.load(open("gpurun_out/${tag}_kernels_1080p_dense%s.json" % name))
    f = d["fused_levels"][0]
    print(name or "split+stream2", "L0 us/iter", f["us_per_gn_iter"], "frac", round(f["frac"], 3), "P1", f["phase_us_per_eval"]["P1_residuals"], "P4", f["phase_us_per_eval"]["P4_reduce"], "| L1", d["fused_levels"][1]["us_per_gn_iter"], "| whole solve us/eval", d["solve_profile"]["us_per_eval"], "evals", d["solve_profile"]["evals"])
PY
done
echo "== headline kernel unchanged?"; bash scripts/gpu_abn.sh ""
