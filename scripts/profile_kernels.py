"""Small driver for ncu captures and per-kernel device timing of the linearize path.

    python scripts/profile_kernels.py [--workload kitti|kitti_dense] [--iters 20]

Prints the cudaEvent time of one host-driven linearize (4 kernels) at level 0 with and without an L2 flush,
its algorithmic bytes and the resulting fraction of the measured HBM peak; then runs one
on-device estimate_pose so that `ncu -k regex:k_estimate_pose` has something to capture."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="kitti_dense")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    from bpvo_b200.engine import Context
    w = bench.WORKLOADS[args.workload]
    sc = bench.make_scene(w, 0xB200)
    p = bench.make_params(w)
    ctx = Context(sc.K, sc.baseline, (sc.rows, sc.cols), p)
    a, b = ctx.frame(), ctx.frame()
    i0, d0 = sc.render(0)
    i1, d1 = sc.render(1)
    a.setData(i0, d0); a.setTemplate(); b.setData(i1, d1)
    T = np.eye(4, dtype=np.float32)
    peak, how = bench.peaks()
    out = {"workload": w["name"], "peak_gbs": peak, "peak_source": how, "levels": []}
    for l in range(p.numPyramidLevels):
        N = a.numPoints(l)
        r, c = a.level_size(l)
        B = bench.algorithmic_bytes_per_iter(N, ctx.channels, r, c)
        ms_cold = ctx.time_linearize(a, b, l, T, iters=args.iters, flush_l2=True)
        ms_warm = ctx.time_linearize(a, b, l, T, iters=args.iters, flush_l2=False)
        out["levels"].append({"level": l, "N": N, "algorithmic_bytes": B, "ms_cold_l2": ms_cold, "ms_warm_l2": ms_warm,
                              "gbs_cold": B / ms_cold / 1e6, "gbs_warm": B / ms_warm / 1e6,
                              "frac_cold": B / ms_cold / 1e6 / peak, "frac_warm": B / ms_warm / 1e6 / peak})
    Tg, stats, evals = ctx.estimatePose(a, b, T)
    out["estimate_pose_evals"] = evals
    # in-kernel phase profile of the persistent solve (CTA 0 cycle counters)
    ctx.set_profiling(True)
    ctx.phase_cycles(reset=True)
    ctx.level_phase_cycles(reset=True)
    ctx.reset_counters()
    tot_ev = 0
    lvl_ev = [0] * p.numPyramidLevels
    lvl_us = [0.0] * p.numPyramidLevels
    for _ in range(5):
        Tg, stats, evals = ctx.estimatePose(a, b, T)
        tot_ev += evals
        for l, (e, u) in enumerate(zip(ctx.last_level_evals(), ctx.last_level_us())):
            lvl_ev[l] += e; lvl_us[l] += u
    cyc = ctx.phase_cycles()
    # the persistent (fused) solve level by level: device time per GN iteration, algorithmic bytes, fraction of the HBM peak
    lpc = ctx.level_phase_cycles()
    out["fused_levels"] = []
    for l in range(p.numPyramidLevels):
        N = a.numPoints(l); r, c = a.level_size(l)
        B = bench.algorithmic_bytes_per_iter(N, ctx.channels, r, c)
        us = lvl_us[l] / max(lvl_ev[l], 1)
        tot = float(sum(lpc[l].values())) or 1.0
        out["fused_levels"].append({"level": l, "N": N, "evals": lvl_ev[l], "us_per_gn_iter": round(us, 3), "algorithmic_bytes": B,
                                    "gbs": B / us / 1e3 if us > 0 else 0.0, "frac": B / us / 1e3 / peak if us > 0 else 0.0,
                                    "phase_us_per_eval": {k: round(us * v / tot, 3) for k, v in lpc[l].items()}})
    ms = ctx.counters()["ms_linearize"]
    fine = cyc.pop("_fine")
    hits, ests = cyc.pop("_bracket_hits"), cyc.pop("_scale_estimates")
    out["bracket_hit_rate"] = hits / max(ests, 1)
    out["bracket_overflows"], out["bracket_misses"] = cyc.pop("_bracket_overflows"), cyc.pop("_bracket_misses")
    tot = float(sum(cyc.values())) or 1.0
    out["solve_profile"] = {"evals": tot_ev, "ms": ms, "us_per_eval": 1e3 * ms / max(tot_ev, 1),
                            "phase_share": {k: round(v / tot, 4) for k, v in cyc.items()},
                            "phase_us_per_eval": {k: round(1e3 * ms * (v / tot) / max(tot_ev, 1), 3) for k, v in cyc.items()},
                            "level_evals": ctx.last_level_evals()}
    # fine marks (only in the -DBP_FINE_PROFILE build, BPVO_B200_LIB=.../libbpvo_b200_fine.so): slot 16 + index
    names = {0: "P1:zero+sync", 1: "P1:loop", 2: "P1:sync", 3: "P1:flush", 4: "P4:loop", 5: "P4:sync", 6: "P4:butterfly", 7: "P4:sync2",
             8: "P4:cta-sum", 9: "pre-solve", 10: "LDLT", 12: "update+sync", 13: "convergence-tests", 16: "barrier1", 17: "select:load-candidates+counters",
             19: "select:binning", 20: "select:find2", 21: "select:list", 22: "select:rank", 23: "next-bracket",
             28: "list:enter", 29: "list:match", 30: "list:append", 24: "exchange:leader-entry", 25: "exchange:leader-gather+post", 26: "exchange:stage2-gather", 27: "exchange:cta-sum"}
    if sum(fine):
        out["fine_us_per_eval"] = {names.get(i, f"slot{16 + i}"): round(f / 1.965e3 / max(tot_ev, 1), 3) for i, f in enumerate(fine) if f}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
