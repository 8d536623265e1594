#!/bin/bash
# round 2, call H: TMA bit-planes kernel again (descriptors in global memory): sanitizer, parity suite, A/B under ncu
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TAG=${1:-r2h}
timeout 300 compute-sanitizer --tool memcheck python scripts/tma_debug.py > gpurun_out/${TAG}_tma_sanitizer.log 2>&1
tail -5 gpurun_out/${TAG}_tma_sanitizer.log
if grep -q "descriptor ok" gpurun_out/${TAG}_tma_sanitizer.log && grep -q "ERROR SUMMARY: 0 errors" gpurun_out/${TAG}_tma_sanitizer.log; then
  echo "TMA kernel clean"
else
  echo "TMA kernel still faulting: falling back to for the rest"; true
fi
timeout 2400 python -m pytest tests -m gpu -q -rs > gpurun_out/${TAG}_pytest_gpu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:bitplanes -c 48 --csv --log-file gpurun_out/${TAG}_bitplanes_tma.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-dense --no-throughput > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:bitplanes -c 48 --csv --log-file gpurun_out/${TAG}_bitplanes_plain.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-dense --no-throughput > /dev/null 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-dense --no-throughput --workload kitti_cfg > gpurun_out/${TAG}_bench_cfg_tma.json 2> gpurun_out/${TAG}_bench_cfg_tma.err
timeout 300 python bench.py --no-cpu-baseline --no-dense --no-throughput --workload kitti_cfg > gpurun_out/${TAG}_bench_cfg_plain.json 2> gpurun_out/${TAG}_bench_cfg_plain.err
timeout 300 python scripts/profile_kernels.py --workload 1080p_dense > gpurun_out/${TAG}_kernels_1080p_dense.json 2> gpurun_out/${TAG}_kernels_1080p_dense.err
timeout 300 python scripts/profile_kernels.py --workload kitti_dense > gpurun_out/${TAG}_kernels_dense.json 2> gpurun_out/${TAG}_kernels_dense.err
grep -E "passed|failed|FAILED|^E  |SKIPPED" gpurun_out/${TAG}_pytest_gpu.log | tail -20
python - <<PY
import json, csv
for v in ("tma","plain"):
    try:
        rows=[r for r in csv.reader(open("gpurun_out/${TAG}_bitplanes_%s.csv"%v)) if len(r)>10 and r[0].isdigit()]
    except Exception as e:
        print(v, e); continue
    agg={}
    for r in rows:
        name=r[4][:28]; grid=r[8]; metric=r[-3]
        try: val=float(r[-1].replace(",",""))
        except: continue
        agg.setdefault((name,grid,metric),[]).append(val)
    for k,vals in sorted(agg.items()): print(v, k, "n=%d"%len(vals), "median=%.1f"%sorted(vals)[len(vals)//2])
for v in ("tma","plain"):
    d=json.load(open("gpurun_out/${TAG}_bench_cfg_%s.json"%v)); print("kitti_cfg", v, round(d["value"],1), round(d["e2e"]["value"],1), d["roofline"]["phase_ms_per_frame"])
for f in ("kernels_1080p_dense","kernels_dense"):
    d=json.load(open("gpurun_out/${TAG}_%s.json"%f))
    print(f, "hit", round(d["bracket_hit_rate"],3), [(L["level"], L["N"], L["us_per_gn_iter"], round(L["frac"],3)) for L in d["fused_levels"]])
PY
