#!/bin/bash
# round 2, call N: whole GPU suite after the 3- / 5-channel descriptors went in; short bench for regressions
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TAG=${1:-r2n}
timeout 3000 python -m pytest tests -m gpu -q -rs > gpurun_out/${TAG}_pytest_gpu.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
timeout 400 python bench.py --no-cpu-baseline --no-dense --no-throughput > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
grep -E "passed|failed|FAILED|^E  |SKIPPED" gpurun_out/${TAG}_pytest_gpu.log | tail -30; tail -2 gpurun_out/${TAG}_smoke.log
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json")); print("bench", round(d["value"],1), round(d["e2e"]["value"],1), d["gn_iters_per_frame"], [ (l["level"], round(l["us_per_gn_iter"],2), round(l["frac"],3)) for l in d["roofline"]["per_level"]])
PY
