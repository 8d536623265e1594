#!/bin/bash
# round 2, call F: TMA bit-planes kernel (parity + A/B under ncu), L2 prefetch A/B on the streaming levels, full suite
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TAG=${1:-r2f}
timeout 2400 python -m pytest tests -m gpu -q -rs > gpurun_out/${TAG}_pytest_gpu.log 2>&1
# descriptor kernel: device time per launch, TMA vs plain loads/stores (ncu serialises and times every launch; compare the two lists)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:bitplanes -c 64 --csv --log-file gpurun_out/${TAG}_bitplanes_tma.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-dense --no-throughput > /dev/null 2>&1
BPVO_B200_NO_TMA=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:bitplanes -c 64 --csv --log-file gpurun_out/${TAG}_bitplanes_plain.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-dense --no-throughput > /dev/null 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-dense --no-throughput --workload kitti_cfg > gpurun_out/${TAG}_bench_cfg_tma.json 2> gpurun_out/${TAG}_bench_cfg_tma.err
BPVO_B200_NO_TMA=1 timeout 300 python bench.py --no-cpu-baseline --no-dense --no-throughput --workload kitti_cfg > gpurun_out/${TAG}_bench_cfg_plain.json 2> gpurun_out/${TAG}_bench_cfg_plain.err
for v in "" _noprefetch; do
  L=$PWD/bpvo_b200/libbpvo_b200$v.so
  BPVO_B200_LIB=$L timeout 300 python scripts/profile_kernels.py --workload 1080p_dense > gpurun_out/${TAG}_kernels_1080p_dense$v.json 2> gpurun_out/${TAG}_kernels_1080p_dense$v.err
  BPVO_B200_LIB=$L timeout 300 python scripts/profile_kernels.py --workload kitti_dense > gpurun_out/${TAG}_kernels_dense$v.json 2> gpurun_out/${TAG}_kernels_dense$v.err
done
grep -E "passed|failed|FAILED|^E  |SKIPPED" gpurun_out/${TAG}_pytest_gpu.log | tail -20
python - <<PY
import json, csv
for v in ("tma","plain"):
    rows=[r for r in csv.reader(open("gpurun_out/${TAG}_bitplanes_%s.csv"%v)) if len(r)>10 and r[0].isdigit()]
    # columns: ID,...,Kernel Name,...,Metric Name,Metric Unit,Metric Value
    agg={}
    for r in rows:
        name=r[4][:40]; metric=r[-3]; val=float(r[-1].replace(",",""))
        agg.setdefault((name,metric),[]).append(val)
    for k,vals in sorted(agg.items()): print(v, k, "n=%d"%len(vals), "max=%.1f min=%.1f"%(max(vals),min(vals)))
for v in ("tma","plain"):
    d=json.load(open("gpurun_out/${TAG}_bench_cfg_%s.json"%v)); print("kitti_cfg", v, d["value"], d["e2e"]["value"], d["roofline"]["phase_ms_per_frame"])
for f in ("kernels_1080p_dense","kernels_1080p_dense_noprefetch","kernels_dense","kernels_dense_noprefetch"):
    d=json.load(open("gpurun_out/${TAG}_%s.json"%f))
    print(f, "hit", round(d["bracket_hit_rate"],3), [(L["level"], L["N"], L["us_per_gn_iter"], round(L["frac"],3)) for L in d["fused_levels"]])
PY
