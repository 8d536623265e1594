#!/usr/bin/env python
"""Per-iteration trace of the Gauss-Newton loop, device loop vs the oracle, on the bench stream.

    python scripts/iter_trace.py [--workload kitti] [--frames 8] [--flags 0] > gpurun_out/iter_trace.json

For every frame: per-level (numIterations, status) of both, and for frame `--dump` the whole table
{level, eval, f_norm, |dp|, max|G|, sigma, ...} of both side by side (what PoseEstimatorBase::run prints at
verbosity kIteration, pose_estimator_base.h:231-247).  Diagnosis tool for iteration-count differences."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="kitti")
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--dump", type=int, default=2)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--ptol", type=float, default=None)
    ap.add_argument("--ftol", type=float, default=None)
    args = ap.parse_args()
    sys.argv = sys.argv[:1]
    import bench
    from bpvo_b200.engine import Context
    from oracle import pyoracle as po
    w = bench.WORKLOADS[args.workload]
    sc = bench.make_scene(w, 0xB200)
    p = bench.make_params(w)
    if args.ptol is not None:
        p.parameterTolerance = args.ptol
    if args.ftol is not None:
        p.functionTolerance = args.ftol
    ctx = Context(sc.K, sc.baseline, (sc.rows, sc.cols), p, flags=args.flags)
    ctx.set_trace(True)
    oest = po.Estimator(ctx.params)
    oest.set_trace(True)
    T0 = np.eye(4, dtype=np.float32)
    out = {"workload": w["name"], "frames": []}
    i0, d0 = sc.render(0)
    gref = ctx.frame(); gref.setData(i0, d0); gref.setTemplate()
    oref = po.Frame(sc.K, sc.baseline, sc.rows, sc.cols, ctx.params, use_rcp=1); oref.set_data(i0, d0); oref.set_template()
    tot_g = tot_o = 0
    for k in range(1, args.frames + 1):
        i1, d1 = sc.render(k)
        gcur = ctx.frame(); gcur.setData(i1, d1)
        ocur = po.Frame(sc.K, sc.baseline, sc.rows, sc.cols, ctx.params, use_rcp=1); ocur.set_data(i1, d1)
        oest.set_trace(True)
        To, so, no = oest.estimate_pose(oref, ocur, T0)
        tr_o = oest.get_trace()
        Tg, sg, ng = ctx.estimatePose(gref, gcur, T0)
        tr_g = ctx.get_trace()
        tot_g += ng; tot_o += no
        rec = {"frame": k, "evals_gpu": ng, "evals_oracle": no,
               "gpu": [(s.numIterations, hex(s.status)) for s in sg], "oracle": [(s["numIterations"], hex(s["status"])) for s in so],
               "pose_rel_err": float(np.abs(Tg - To).max() / np.abs(To).max())}
        if k == args.dump:
            rec["trace_gpu"] = tr_g.tolist()
            rec["trace_oracle"] = tr_o.tolist()
        out["frames"].append(rec)
        gcur.close()
    out["evals_per_frame_gpu"] = tot_g / args.frames
    out["evals_per_frame_oracle"] = tot_o / args.frames
    print(json.dumps(out))


if __name__ == "__main__":
    main()
