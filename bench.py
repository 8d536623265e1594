#!/usr/bin/env python
"""bench.py -- frames/sec and GN-iterations/sec of the per-frame Gauss-Newton dense-alignment path
(bpvo's VisualOdometry::addFrame) on synthetic KITTI-sized bit-planes streams.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload kitti|vga|kitti_dense|1080p|1080p_dense]

A "step" is one addFrame() = image+disparity in -> pose out (pyramid, bit-planes descriptors, the whole
coarse-to-fine GN solve, key-frame work when it triggers).  Prints ONE JSON line (rank 0).

  value : frames/s with the step's inputs already resident in HBM when the timed region starts
  e2e   : frames/s through the same public call with PINNED HOST buffers (H2D inside the timed region)
  roofline : the dominant kernel (the persistent on-device GN solve) against the measured HBM peak
  cpu_baseline : the CPU restatement of the reference (oracle/) timed on this box's host cores

--impl reference times the reference's own CPU implementation of the path: bpvo's own sources compiled from
/root/reference against stand-in Eigen/OpenCV headers (oracle/_ref/libbpvo_ref.so, prebuilt here and shipped to
the GPU box), falling back to the oracle port if that library is absent.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2]: KITTI 1241x376 bit-planes (8 ch), 4 levels, Tukey IRLS
    "kitti": dict(scene="kitti", descriptor="bitplanes", levels=4, loss="tukey", nms=1,
                  name="kitti_1241x376_bitplanes_8ch_4levels_tukey"),
    "kitti_dense": dict(scene="kitti", descriptor="bitplanes", levels=4, loss="tukey", nms=-1,
                        name="kitti_1241x376_bitplanes_8ch_4levels_tukey_dense"),
    # the same stream at the tolerances / loss the reference SHIPS for KITTI (conf/kitti_bitplanes.cfg: 5 levels, Huber, pTol 1e-6,
    # fTol 1e-4, maxIterations 100, key-frame thresholds 1.0 m / 2.5 / 0.6); minSaliency stays at the ctor default 0.1: the
    # file's 2.5 selects no point at all with the reference's bit-planes saliency (channel 0 only, max 2)
    "kitti_cfg": dict(scene="kitti", descriptor="bitplanes", levels=5, loss="huber", nms=1,
                      extra=dict(parameterTolerance=1e-6, functionTolerance=1e-4, maxIterations=100, minTranslationMagToKeyFrame=1.0,
                                 minRotationMagToKeyFrame=2.5, maxFractionOfGoodPointsToKeyFrame=0.6),
                      name="kitti_1241x376_bitplanes_8ch_5levels_huber_shipped_tolerances"),
    "vga": dict(scene="vga", descriptor="intensity", levels=4, loss="huber", nms=1,
                name="vga_640x480_intensity_4levels_huber"),
    "1080p": dict(scene="1080p", descriptor="bitplanes", levels=5, loss="tukey", nms=1,
                  name="1080p_bitplanes_8ch_5levels_tukey"),
    # BASELINE.json configs[3]: 1920x1080 bit-planes, 5 levels, Tukey, template points sharded across the GPUs (dense selection)
    "1080p_dense": dict(scene="1080p", descriptor="bitplanes", levels=5, loss="tukey", nms=-1,
                        name="1080p_bitplanes_8ch_5levels_tukey_dense"),
}


def make_scene(w, seed):
    from bpvo_b200 import synth
    return {"kitti": synth.scene_kitti, "vga": synth.scene_vga, "1080p": synth.scene_1080p}[w["scene"]](seed=seed)


def make_params(w):
    from bpvo_b200.types import AlgorithmParameters, DescriptorType, LossFunctionType, VerbosityType
    return AlgorithmParameters(
        descriptor={"intensity": DescriptorType.kIntensity, "bitplanes": DescriptorType.kBitPlanes}[w["descriptor"]],
        numPyramidLevels=w["levels"],
        lossFunction={"tukey": LossFunctionType.kTukey, "huber": LossFunctionType.kHuber, "l2": LossFunctionType.kL2}[w["loss"]],
        nonMaxSuppRadius=w["nms"], verbosity=VerbosityType.kSilent, **w.get("extra", {}))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons DURING the timed region"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def algorithmic_bytes_per_iter(N, C, rows, cols):
    """points + (J + I0) + descriptor footprint touched + (r + w written), reference data model"""
    return 16 * N + (24 + 4) * N * C + min(16 * N * C, 4 * C * rows * cols) + 8 * N * C


# --------------------------------------------------------------------------------------------------
def _cpu_vo(w, sc, p, nthreads):
    """-> (vo object with add_frame_raw(img_ptr, disp_ptr, OrcResult), kind, note).  Prefers the REAL reference
    (oracle/_ref: bpvo's own sources compiled against stand-in Eigen/OpenCV headers); falls back to the oracle port."""
    from oracle import pyoracle as po
    L = po.ref_lib()
    if L is not None:
        L.ref_set_num_threads(int(nthreads))
        return po.RefVisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p), "reference", \
            "bpvo's own sources (vo.cc ... rigid_body_warp.cc) compiled -O3 -msse4.1 -mavx WITH_SIMD WITH_OPENMP against stand-in Eigen/OpenCV headers"
    return po.VisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p, use_rcp=1, num_threads=nthreads), "port", \
        "oracle port (oracle/_ref not available on this box)"


def _time_cpu(vo, ptrs, first, count):
    """-> (seconds, linearize evaluations or None).  The reference's Result carries no evaluation count (only the statistics of
    the LAST estimatePose of a frame: a key-frame's first solve is invisible); see _count_evals for the exact number."""
    from oracle import pyoracle as po
    res = po.OrcResult()
    evals, known = 0, True
    t0 = time.perf_counter()
    for k in range(first, first + count):
        vo.add_frame_raw(ptrs[k][0], ptrs[k][1], res)
        known = known and res.numFunEvals > 0
        evals += res.numFunEvals
    return time.perf_counter() - t0, (evals if known else None)


def _count_evals(sc, p, ptrs, first, count):
    """exact GN-iteration count of the reference on frames [first, first + count): the oracle port (bit-identical poses,
    per-level iteration counts and key-frame decisions, tests/test_oracle_vs_reference.py) counts every linearize(); untimed"""
    from oracle import pyoracle as po
    vo = po.VisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p, use_rcp=1, num_threads=max(1, min(os.cpu_count() or 1, 8)))
    res = po.OrcResult()
    evals = 0
    for k in range(first + count):
        vo.add_frame_raw(ptrs[k][0], ptrs[k][1], res)
        if k >= first:
            evals += res.numFunEvals
    return evals


def run_reference(args, w, rank, world):
    """CPU arm: the reference's own implementation of the path on the host cores, rank 0 only.  Thread count: the reference's
    OpenMP parallel_for (descriptor channels, residual channels) is tried with 1 and with min(8, cores) threads on two
    frames each and the faster setting is used for the timed run (on most hosts that is ONE thread: the parallel regions
    are tiny and the scale / weights / normal-equation passes are serial in the reference)."""
    if rank != 0:
        return
    sc = make_scene(w, 0xB200)
    p = make_params(w)
    ncpu = os.cpu_count() or 1
    nframes = args.steps + args.warmup + 1
    frames = [sc.render(k) for k in range(nframes)]
    ptrs = [(f[0].ctypes.data_as(C.POINTER(C.c_uint8)), f[1].ctypes.data_as(C.POINTER(C.c_float))) for f in frames]
    cand = sorted({1, max(1, min(ncpu, 8))})
    best = None
    for nt in cand:
        vo, kind, note = _cpu_vo(w, sc, p, nt)
        _time_cpu(vo, ptrs, 0, 1)
        dt, _ = _time_cpu(vo, ptrs, 1, min(2, nframes - 1))
        if best is None or dt < best[0]:
            best = (dt, nt)
        del vo
    nthreads = best[1]
    vo, kind, note = _cpu_vo(w, sc, p, nthreads)
    _time_cpu(vo, ptrs, 0, args.warmup + 1)
    dt, evals = _time_cpu(vo, ptrs, args.warmup + 1, args.steps)
    if evals is None:
        evals = _count_evals(sc, p, ptrs, args.warmup + 1, args.steps)
    fps = args.steps / dt
    line = {
        "impl": "reference", "metric": "frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (fp64 projection + bilinear blend)", "data": "synthetic",
        "gn_iters_per_sec": evals / dt, "gn_iters_per_frame": evals / float(args.steps),
        "config": {"workload": w["name"], "streams": 1, "rows": sc.rows, "cols": sc.cols, "note": note},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": nthreads, "kind": kind, "host_cores": ncpu,
                         "sample": f"{args.steps} consecutive addFrame calls of the same synthetic stream, {nthreads} thread(s) "
                                   f"(faster of {cand} on this host)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
def timed_stream(vo, ptr_pairs, warmup, steps, dist, torch, sampler=None):
    """W untimed + K timed addFrame calls; returns (device ms, wall ms, GN evals, key-frames) for the K steps."""
    ctx = vo.ctx
    evals = kfs = 0
    for k in range(warmup):
        vo.addFrameRaw(ptr_pairs[k][0], ptr_pairs[k][1], want_cloud=False)
    ctx.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    t0 = time.perf_counter()
    ctx.timer_start()
    for k in range(warmup, warmup + steps):
        r = vo.addFrameRaw(ptr_pairs[k][0], ptr_pairs[k][1], want_cloud=False)
        evals += r.numFunEvals
        kfs += int(r.isKeyFrame)
    ms = ctx.timer_stop()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if sampler else None
    return ms, wall, evals, kfs, clocks


def run_ours(args, w, rank, world, local_rank):
    import torch
    from bpvo_b200 import VisualOdometry
    from bpvo_b200.engine import PinnedBuffer
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    # one independent stream per GPU (BASELINE.json configs[4]: seeds 0xB200 + g), weak scaling, no data-path collective
    sc = make_scene(w, 0xB200 + rank)
    p = make_params(w)
    n_roof = 8
    n_total = args.warmup + args.steps + 1 + n_roof
    frames = [sc.render(k) for k in range(n_total)]
    npx = sc.rows * sc.cols

    # (1) inputs resident in HBM
    d_img = [torch.from_numpy(f[0]).to(dev) for f in frames]
    d_dsp = [torch.from_numpy(f[1]).to(dev) for f in frames]
    torch.cuda.synchronize()
    dev_ptrs = [(a.data_ptr(), b.data_ptr()) for a, b in zip(d_img, d_dsp)]
    # (2) inputs in pinned host memory
    pin_img = PinnedBuffer((n_total, sc.rows, sc.cols), np.uint8)
    pin_dsp = PinnedBuffer((n_total, sc.rows, sc.cols), np.float32)
    for k, f in enumerate(frames):
        pin_img.array[k] = f[0]
        pin_dsp.array[k] = f[1]
    host_ptrs = [(pin_img.ptr + k * npx, pin_dsp.ptr + k * npx * 4) for k in range(n_total)]

    def fresh_vo():
        v = VisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p, device_id=local_rank)
        v.addFrameRaw(host_ptrs[0][0], host_ptrs[0][1], want_cloud=False)    # frame 0 = first key-frame (template only)
        return v

    # ---- device-resident run -> `value` -----------------------------------------------------------
    vo = fresh_vo()
    vo.ctx.reset_counters()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, wall, evals, kfs, clocks = timed_stream(vo, dev_ptrs[1:], args.warmup, args.steps, dist, torch, sampler)
    cnt = vo.ctx.counters()
    launches_total = cnt["launches"]
    launches_timed = int(round(launches_total * args.steps / float(args.steps + args.warmup)))
    vo.close()

    # ---- host-buffer run -> `e2e` -----------------------------------------------------------------
    vo = fresh_vo()
    vo.ctx.reset_counters()
    ms_e, wall_e, evals_e, kfs_e, _ = timed_stream(vo, host_ptrs[1:], args.warmup, args.steps, dist, torch)
    cnt_e = vo.ctx.counters()
    h2d = cnt_e["h2d_bytes"] / float(args.steps + args.warmup)
    d2h = cnt_e["d2h_bytes"] / float(args.steps + args.warmup)

    # ---- the same from PAGEABLE host memory (plain numpy arrays): two extra host memcpys per frame into the engine's pinned staging ----
    pageable_fps = None
    if world == 1:
        vo_p = fresh_vo()
        pg_ptrs = [(f[0].ctypes.data, f[1].ctypes.data) for f in frames]
        ms_p, _, _, _, _ = timed_stream(vo_p, pg_ptrs[1:], args.warmup, args.steps, dist, torch)
        pageable_fps = args.steps / (ms_p * 1e-3)
        vo_p.close()

    # ---- roofline of the dominant kernel (persistent GN solve), measured live with CUDA events -----
    roof = None
    if rank == 0:
        ctx = vo.ctx
        ctx.set_profiling(True)
        ctx.reset_counters()
        ref = vo.ref_frame()
        alg_bytes, ms_lin, n_solves = 0.0, 0.0, 0
        lvl_us = [0.0] * p.numPyramidLevels
        lvl_ev = [0] * p.numPyramidLevels
        lvl_bytes = [0.0] * p.numPyramidLevels
        phase = {k2: 0.0 for k2 in ("ms_upload", "ms_pyramid", "ms_descriptor", "ms_template", "ms_linearize")}
        Cch = ctx.channels
        sizes = [ref.level_size(l) for l in range(p.numPyramidLevels)]
        for k in range(args.warmup + args.steps, args.warmup + args.steps + n_roof):
            npts = [vo.ref_frame().numPoints(l) for l in range(p.numPyramidLevels)]    # template the solve runs against
            c0 = ctx.counters()
            r = vo.addFrameRaw(host_ptrs[1 + k][0], host_ptrs[1 + k][1], want_cloud=False)
            c1 = ctx.counters()
            ev = ctx.last_level_evals()
            for k2 in phase:
                phase[k2] += c1[k2] - c0[k2]
            if r.isKeyFrame:
                continue      # key-frames run two solves against two templates; keep the byte accounting exact by skipping them
            ms_lin += c1["ms_linearize"] - c0["ms_linearize"]
            n_solves += c1["solve_calls"] - c0["solve_calls"]
            us = ctx.last_level_us()
            for l in range(p.numPyramidLevels):
                b = ev[l] * algorithmic_bytes_per_iter(npts[l], Cch, sizes[l][0], sizes[l][1])
                alg_bytes += b
                lvl_us[l] += us[l]; lvl_ev[l] += ev[l]; lvl_bytes[l] += b
        n_solves = max(1, n_solves)
        peak, how = peaks()
        achieved = alg_bytes / (ms_lin * 1e-3) / 1e9 if ms_lin > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")      # dram__bytes_{read,write}.sum of one `ncu --set full` capture of this kernel
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = float(tj["dram_bytes_read"]) + float(tj["dram_bytes_write"])
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "kernel": "k_estimate_pose<8> (persistent on-device GN solve, all levels/iterations in one launch)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_note": "bytes per launch from profiles/ncu_traffic.json: far BELOW the algorithmic bytes because the working set is L2-resident",
                "peak_source": how, "ms_per_launch": ms_lin / n_solves,
                "algorithmic_bytes_per_launch": alg_bytes / n_solves,
                "per_level": [{"level": l, "gn_iters": lvl_ev[l], "us_per_gn_iter": lvl_us[l] / max(lvl_ev[l], 1),
                               "achieved": lvl_bytes[l] / max(lvl_us[l], 1e-9) / 1e3, "frac": lvl_bytes[l] / max(lvl_us[l], 1e-9) / 1e3 / peak}
                              for l in range(p.numPyramidLevels)],
                "note": "KITTI semi-dense levels (5-30 k points, 13 MB per iteration, L2 resident) are latency / synchronisation bound, not HBM bound; "
                        "the same kernel on HBM-bound levels: roofline_dense_variant / roofline_hbm_bound_variant (per level, in-kernel global timer)",
                "phase_ms_per_frame": {k2: v / float(n_roof) for k2, v in phase.items()}}
        ctx.set_profiling(False)
    vo.close()

    # ---- reduce over ranks: max time, total work ----------------------------------------------------
    t = torch.tensor([ms, ms_e, float(evals), float(evals_e)], dtype=torch.float64, device=dev)
    if dist is not None:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, ms_e = float(tmax[0]), float(tmax[1])
        evals, evals_e = float(tsum[2]), float(tsum[3])
    frames_total = args.steps * world
    value = frames_total / (ms * 1e-3)
    e2e = frames_total / (ms_e * 1e-3)

    cpu = None
    dense = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(w, args)
    hbm_bound = None
    if rank == 0 and world == 1 and args.workload == "kitti" and not args.no_dense:
        dense = fused_levels_roofline("kitti_dense", local_rank)
        hbm_bound = hbm_bound_variant(local_rank)
        hbm_bound["fused_kernel"] = fused_levels_roofline("1080p_dense", local_rank, solves=2)

    throughput = None
    if rank == 0 and world == 1 and not args.no_throughput:
        throughput = throughput_variant(sc, p, host_ptrs, local_rank, torch, args.warmup, min(args.steps, 48))

    stereo = None
    if rank == 0 and world == 1 and args.workload == "kitti" and not args.no_stereo:
        stereo = stereo_variant(sc, p, local_rank, torch, args)

    sharded = None
    if world > 1 and args.workload == "kitti" and not args.no_dense:
        sharded = sharded_variant(rank, world, local_rank, dist, torch)

    if rank == 0:
        line = {
            "metric": "frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (fp64 projection + bilinear blend, as the reference)", "data": "synthetic",
            "gn_iters_per_sec": evals / (ms * 1e-3), "gn_iters_per_frame": evals / frames_total, "keyframes_in_timed_region": kfs,
            "config": {"workload": w["name"], "rows": sc.rows, "cols": sc.cols, "streams": world, "parallelism": f"replicas x{world} (one independent stream per GPU)",
                       "l2_policy": "each step consumes a NEW frame (2.3 MB input, 20 MB of fresh descriptors); working set is L2-resident by nature, no artificial flush",
                       "timing": "cudaEvents on the engine stream around K addFrame calls, max over ranks", "wall_ms": wall},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "gn_iters_per_sec": evals_e / (ms_e * 1e-3), "ms_per_step": ms_e / args.steps,
                    "from_pageable_host_memory": pageable_fps,
                    "note": "value: inputs in PINNED host memory (image on the engine stream, disparity map on a copy stream beside the frame's kernels); "
                            "from_pageable_host_memory: plain malloc'ed arrays, staged through the engine's pinned buffers (two extra host memcpys per frame)"},
            "gpu_launches": launches_timed,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "roofline_dense_variant": dense,
            "roofline_hbm_bound_variant": hbm_bound,
            "throughput_mode": throughput,
            "sharded_1080p_dense_variant": sharded,
            "upstream_stereo_variant": stereo,
        }
        emit(line)
    pin_img.free(); pin_dsp.free()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def stereo_variant(sc, p, local_rank, torch, args, pairs=24):
    """SURVEY 8(f) N4, beside the headline: the disparity producer in front of the path (OpenCV StereoBM with conf/kitti.cfg's
    settings: BlockMatching, SADWindowSize 9, 128 disparities) on the GPU, and image PAIRS -> poses through addStereoFrame.
    CPU side: the oracle port on all host threads, and cv2's own StereoBM (the library the reference calls) when importable."""
    import time
    from bpvo_b200 import VisualOdometry
    from bpvo_b200.engine import PinnedBuffer
    from bpvo_b200.stereo import StereoAlgorithm
    nd, wsz = 128, 9
    npx = sc.rows * sc.cols
    n = pairs + 4
    pinL = PinnedBuffer((n, sc.rows, sc.cols), np.uint8); pinR = PinnedBuffer((n, sc.rows, sc.cols), np.uint8)
    pinD = PinnedBuffer((sc.rows, sc.cols), np.float32)
    for k in range(n):
        pinL.array[k] = sc.render(k)[0]; pinR.array[k] = sc.render_right(k)
    st = StereoAlgorithm((sc.rows, sc.cols), device_id=local_rank, numberOfDisparities=nd, SADWindowSize=wsz)
    dL = torch.from_numpy(pinL.array).cuda(local_rank); dR = torch.from_numpy(pinR.array).cuda(local_rank)
    dD = torch.empty((sc.rows, sc.cols), dtype=torch.float32, device=dL.device)
    torch.cuda.synchronize()
    kern_ms = []
    for k in range(n):                                 # resident in HBM, result stays in HBM: device time of the three kernels
        st.run_raw(dL[k].data_ptr(), dR[k].data_ptr(), dD.data_ptr())
        if k >= 4:
            kern_ms.append(st.last_kernel_ms())
    t0 = time.perf_counter()
    for k in range(4, n):                              # pinned host buffers in, pinned host disparity map out
        st.run_raw(pinL.ptr + k * npx, pinR.ptr + k * npx, pinD.ptr)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / pairs
    d_gpu = pinD.array.copy()
    valid = float((d_gpu >= 0).mean())
    truth = sc.render(n - 1)[1]
    err = float(np.abs(d_gpu - truth)[d_gpu >= 0].mean())
    # pairs -> poses
    vo = VisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p, device_id=local_rank)
    for k in range(4):
        vo._lib.bpvo_b200_vo_add_stereo_frame(vo.h, st.h, pinL.ptr + k * npx, pinR.ptr + k * npx, __import__("ctypes").byref(vo._res))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    evals = 0
    for k in range(4, n):
        rc = vo._lib.bpvo_b200_vo_add_stereo_frame(vo.h, st.h, pinL.ptr + k * npx, pinR.ptr + k * npx, __import__("ctypes").byref(vo._res))
        assert rc == 0
        evals += vo._res.numFunEvals
    pipe_ms = (time.perf_counter() - t0) * 1e3 / pairs
    vo.close()
    # CPU: oracle port (OpenMP over rows), and cv2 itself if it is there
    from oracle import pyoracle as po
    L0, R0 = pinL.array[n - 1], pinR.array[n - 1]
    t0 = time.perf_counter(); reps = 3
    for _ in range(reps):
        want16, wantf = po.stereo_bm(L0, R0, nd, wsz)
    cpu_ms = (time.perf_counter() - t0) * 1e3 / reps
    cv2_ms, cv2_threads, cv2_equal = None, None, None
    try:
        import cv2
        bm = cv2.StereoBM_create(nd, wsz)
        bm.compute(L0, R0)
        t0 = time.perf_counter()
        for _ in range(10):
            ref16 = bm.compute(L0, R0)
        cv2_ms = (time.perf_counter() - t0) * 1e3 / 10
        cv2_threads = cv2.getNumThreads()
        cv2_equal = bool(np.array_equal(ref16, want16))
    except Exception:       # noqa: BLE001
        pass
    st.close(); pinL.free(); pinR.free(); pinD.free()
    k_ms = float(np.median(kern_ms))
    peak, how = peaks()
    alg_bytes = 2.0 * npx + 4.0 * npx                        # two u8 images in, the f32 map out
    cells = (sc.rows - wsz + 1) * (sc.cols - nd + 1 - wsz + 1) * nd
    return {"what": "OpenCV StereoBM (conf/kitti.cfg: BlockMatching, SADWindowSize 9, numberOfDisparities 128) + pairs -> poses; bit-exact vs the oracle in tests/test_gpu_stereo.py",
            "pairs_per_sec_resident": 1e3 / k_ms, "kernel_ms_per_pair": k_ms, "kernels_per_pair": 3,
            "pairs_per_sec_e2e_pinned_host": 1e3 / e2e_ms, "h2d_bytes_per_pair": 2 * npx, "d2h_bytes_per_pair": 4 * npx,
            "identical_to_oracle": bool(np.array_equal(d_gpu, wantf)), "valid_fraction": valid, "mean_abs_error_vs_rendered_disparity_px": err,
            "pairs_to_poses": {"frames_per_sec": 1e3 / pipe_ms, "ms_per_frame": pipe_ms, "gn_iters_per_frame": evals / float(pairs),
                               "note": "bpvo_b200_vo_add_stereo_frame from pinned host pairs: block matching + addFrame, disparity never leaves the device"},
            "roofline": {"bound": "alu / shared-memory loads (integer SAD), not hbm", "algorithmic_bytes": alg_bytes, "hbm_frac": alg_bytes / (k_ms * 1e-3) / 1e9 / peak,
                         "window_sad_cells": cells, "giga_cells_per_sec": cells / (k_ms * 1e-3) / 1e9, "peak_source": how},
            "cpu_baseline": {"oracle_port_ms": cpu_ms, "oracle_threads": os.cpu_count(), "cv2_ms": cv2_ms, "cv2_threads": cv2_threads, "cv2_identical_to_oracle": cv2_equal,
                             "kind": "port (+ the third-party library itself when importable)"}}


def throughput_variant(sc, p, host_ptrs, local_rank, torch, warmup, steps, streams_list=(2, 4, 8)):
    """Throughput mode beside the single-stream headline: S independent VisualOdometry streams on ONE GPU, one host thread
    each, the on-device GN loop of every stream confined to 148 / S SMs (bpvo_b200_set_solver_ctas) so that the solves run
    side by side.  Frames come from pinned host memory (H2D inside the timed region); wall clock between two thread
    barriers with a device synchronize inside, i.e. frames/s of the whole GPU."""
    import threading
    from bpvo_b200 import VisualOdometry
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    out = []
    for S in streams_list:
        ctas = max(1, sms // S)
        vos = []
        for _ in range(S):
            v = VisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p, device_id=local_rank)
            v.ctx.set_solver_ctas(ctas)
            v.addFrameRaw(host_ptrs[0][0], host_ptrs[0][1], want_cloud=False)
            vos.append(v)
        gate = threading.Barrier(S + 1)
        evals = [0] * S
        errs = []

        def worker(i):
            try:
                v = vos[i]
                for k in range(1, 1 + warmup):
                    v.addFrameRaw(host_ptrs[k][0], host_ptrs[k][1], want_cloud=False)
                v.ctx.synchronize()
                gate.wait()
                for k in range(1 + warmup, 1 + warmup + steps):
                    evals[i] += v.addFrameRaw(host_ptrs[k][0], host_ptrs[k][1], want_cloud=False).numFunEvals
                v.ctx.synchronize()
            except Exception as e:      # noqa: BLE001
                errs.append(repr(e))
                gate.abort()
                return
            gate.wait()

        th = [threading.Thread(target=worker, args=(i,)) for i in range(S)]
        for t in th:
            t.start()
        try:
            gate.wait()
            t0 = time.perf_counter()
            gate.wait()
            dt = time.perf_counter() - t0
        except threading.BrokenBarrierError:
            dt = None
        for t in th:
            t.join()
        for v in vos:
            v.close()
        if dt is None or errs:
            out.append({"streams_per_gpu": S, "error": errs[:1]})
            continue
        out.append({"streams_per_gpu": S, "ctas_per_stream": ctas, "value": S * steps / dt, "unit": "frames/s",
                    "gn_iters_per_sec": sum(evals) / dt, "ms_per_frame_per_stream": 1e3 * dt / steps,
                    "timing": "wall clock around S host threads x K addFrame calls from pinned host buffers (H2D inside), device synchronized at both ends"})
    return out


def sharded_variant(rank, world, local_rank, dist, torch):
    """BASELINE.json configs[3] beside the replica numbers (N > 1 only): ONE 1080p dense bit-planes stream whose template
    points are sharded over all N GPUs, the GN loop on the device, exchanges inside the kernel over NVLink (peer-memory
    mode; levels under 131072 points stay replicated).  Reports us per GN iteration against the same solve on one GPU."""
    from bpvo_b200.engine import Context
    w = WORKLOADS["1080p_dense"]
    sc = make_scene(w, 0xB200)
    p = make_params(w)
    i0, d0 = sc.render(0); i1, d1 = sc.render(1)
    T0 = np.eye(4, dtype=np.float32)
    out = {"workload": w["name"], "n_gpus": world}
    poses = {}
    for mode in ("single_gpu", "sharded"):
        ctx = Context(sc.K, sc.baseline, (sc.rows, sc.cols), p, device_id=local_rank)
        if mode == "sharded":
            uid = [Context.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            ctx.comm_init(rank, world, uid[0])
            ctx.peer_init_distributed(dist)
        a, b = ctx.frame(), ctx.frame()
        a.setData(i0, d0); a.setTemplate(); b.setData(i1, d1)
        ctx.estimatePose(a, b, T0)
        ctx.set_profiling(True); ctx.reset_counters()
        dist.barrier()
        evals = 0
        lvl_us = [0.0] * p.numPyramidLevels
        lvl_ev = [0] * p.numPyramidLevels
        for _ in range(3):
            T, stats, n = ctx.estimatePose(a, b, T0)
            evals += n
            for l, (e, u) in enumerate(zip(ctx.last_level_evals(), ctx.last_level_us())):
                lvl_ev[l] += e; lvl_us[l] += u
        poses[mode] = T
        t = torch.tensor([ctx.counters()["ms_linearize"]], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # correctness on the record: every rank must hold the bit-identical pose
        tp = torch.from_numpy(np.ascontiguousarray(T)).to(f"cuda:{local_rank}")
        ref = tp.clone(); dist.broadcast(ref, src=0)
        same = torch.tensor([1 if torch.equal(tp, ref) else 0], device=f"cuda:{local_rank}")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        out[mode] = {"us_per_gn_iter": 1e3 * float(t[0]) / max(evals, 1), "gn_iters": evals, "evals_per_solve": evals / 3.0,
                     "status_per_level": [hex(s.status) for s in stats],
                     "poses_identical_across_ranks": bool(int(same[0])),
                     "us_per_gn_iter_per_level": [round(u / max(e, 1), 2) for u, e in zip(lvl_us, lvl_ev)],
                     "points_per_level_local": [a.numPoints(l) for l in range(p.numPyramidLevels)]}
        if mode == "sharded":
            ctx.comm_destroy()
        a.close(); b.close(); ctx.close()
    d = np.abs(poses["sharded"].astype(np.float64) - poses["single_gpu"].astype(np.float64)).max()
    out["pose_rel_err_vs_single_gpu"] = float(d / np.abs(poses["single_gpu"]).max())
    gt = np.array(sc.relative_pose(0, 1))
    out["translation_err_vs_ground_truth_m"] = float(np.abs(poses["sharded"][:3, 3] - gt[:3, 3]).max())
    out["parity_ok"] = bool(out["pose_rel_err_vs_single_gpu"] < 1e-4 and out["sharded"]["poses_identical_across_ranks"])
    out["speedup_vs_one_gpu"] = out["single_gpu"]["us_per_gn_iter"] / out["sharded"]["us_per_gn_iter"]
    out["sharded_efficiency"] = out["speedup_vs_one_gpu"] / world
    return out


def hbm_bound_variant(local_rank):
    """where the path IS bandwidth bound: one host-driven linearize (4 kernels) at level 0 of the dense 1080p workload
    (BASELINE.json configs[3] on one GPU: 1.83 M points, 622 MB algorithmic per GN iteration, working set beyond the L2),
    cudaEvent time per linearize with an L2 flush (256 MB memset) before every repetition, against the measured HBM peak"""
    from bpvo_b200.engine import Context
    w = WORKLOADS["1080p_dense"]
    sc = make_scene(w, 0xB200)
    p = make_params(w)
    ctx = Context(sc.K, sc.baseline, (sc.rows, sc.cols), p, device_id=local_rank)
    a, b = ctx.frame(), ctx.frame()
    i0, d0 = sc.render(0); i1, d1 = sc.render(1)
    a.setData(i0, d0); a.setTemplate(); b.setData(i1, d1)
    T = np.eye(4, dtype=np.float32)
    N = a.numPoints(0)
    r, c = a.level_size(0)
    B = algorithmic_bytes_per_iter(N, 8, r, c)
    ctx.time_linearize(a, b, 0, T, iters=3, flush_l2=False)
    ms_cold = ctx.time_linearize(a, b, 0, T, iters=10, flush_l2=True)
    ms_warm = ctx.time_linearize(a, b, 0, T, iters=10, flush_l2=False)
    a.close(); b.close(); ctx.close()
    peak, how = peaks()
    return {"workload": w["name"] + " level 0, host-driven linearize (k_residuals + k_select<2,3> + k_reduce)", "bound": "hbm", "points": N,
            "algorithmic_bytes_per_iter": B, "us_per_iter_l2_flushed": 1e3 * ms_cold, "us_per_iter_warm": 1e3 * ms_warm,
            "achieved": B / ms_cold / 1e6, "peak": peak, "unit": "GB/s", "frac": B / ms_cold / 1e6 / peak, "frac_warm": B / ms_warm / 1e6 / peak,
            "peak_source": how}


def fused_levels_roofline(workload, local_rank, solves=3):
    """the persistent (fused) GN kernel on a dense selection, level by level: device time per GN iteration from the kernel's own
    global-timer stamps (LevelStats.us), algorithmic bytes of SURVEY.md 8(d), fraction of the measured HBM peak.  One frame pair,
    `solves` whole coarse-to-fine solves after a warm-up solve."""
    from bpvo_b200.engine import Context
    w = WORKLOADS[workload]
    sc = make_scene(w, 0xB200)
    p = make_params(w)
    ctx = Context(sc.K, sc.baseline, (sc.rows, sc.cols), p, device_id=local_rank)
    a, b = ctx.frame(), ctx.frame()
    i0, d0 = sc.render(0); i1, d1 = sc.render(1)
    a.setData(i0, d0); a.setTemplate(); b.setData(i1, d1)
    T0 = np.eye(4, dtype=np.float32)
    ctx.estimatePose(a, b, T0)
    L = p.numPyramidLevels
    us, ev = [0.0] * L, [0] * L
    for _ in range(solves):
        ctx.estimatePose(a, b, T0)
        for l, (e, u) in enumerate(zip(ctx.last_level_evals(), ctx.last_level_us())):
            ev[l] += e; us[l] += u
    peak, how = peaks()
    levels = []
    for l in range(L):
        N = a.numPoints(l); r, c = a.level_size(l)
        B = algorithmic_bytes_per_iter(N, ctx.channels, r, c)
        t = us[l] / max(ev[l], 1)
        levels.append({"level": l, "points": N, "gn_iters": ev[l], "us_per_gn_iter": t, "algorithmic_bytes_per_iter": B,
                       "achieved": B / t / 1e3 if t > 0 else 0.0, "frac": B / t / 1e3 / peak if t > 0 else 0.0})
    tot_b = sum(lv["algorithmic_bytes_per_iter"] * lv["gn_iters"] for lv in levels)
    tot_us = sum(us)
    a.close(); b.close(); ctx.close()
    best = max(levels, key=lambda lv: lv["frac"])
    return {"workload": w["name"], "kernel": "k_estimate_pose<8> (the fused persistent GN kernel), per pyramid level", "bound": "hbm", "peak": peak, "unit": "GB/s",
            "peak_source": how, "levels": levels, "whole_solve": {"achieved": tot_b / tot_us / 1e3, "frac": tot_b / tot_us / 1e3 / peak, "us_per_gn_iter": tot_us / max(sum(ev), 1)},
            "best_level": best["level"], "achieved": best["achieved"], "frac": best["frac"]}


def cpu_baseline(w, args):
    """the reference's CPU implementation (oracle/_ref, else the oracle port), 1 thread = the reference's default build
    (WITH_TBB off, README recommends a single thread), on a bounded sample of the same stream"""
    sc = make_scene(w, 0xB200)
    p = make_params(w)
    vo, kind, note = _cpu_vo(w, sc, p, 1)
    budget_s = 10.0
    frames = [sc.render(k) for k in range(49)]
    ptrs = [(f[0].ctypes.data_as(C.POINTER(C.c_uint8)), f[1].ctypes.data_as(C.POINTER(C.c_float))) for f in frames]
    _time_cpu(vo, ptrs, 0, 1)
    t_used, n, evals = 0.0, 0, 0
    while t_used < budget_s and n < 48:
        dt, ev = _time_cpu(vo, ptrs, 1 + n, 1)
        t_used += dt; n += 1; evals = None if (ev is None or evals is None) else evals + ev
    if evals is None:
        evals = _count_evals(sc, p, ptrs, 1, n)
    return {"value": n / t_used, "unit": "frames/s", "cores": 1, "kind": kind, "gn_iters_per_sec": evals / t_used, "gn_iters_per_frame": evals / float(n), "note": note,
            "sample": f"first {n} addFrame calls of the same synthetic stream ({t_used:.1f} s of CPU work), single thread = reference default build"}


_REAL_STDOUT = None


def emit(line: dict):
    """the ONE JSON line goes to the process's original stdout; everything else any library prints (NCCL's version banner,
    torchrun notices) was redirected to stderr at start-up"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="kitti", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-selection roofline variant")
    ap.add_argument("--no-throughput", action="store_true", help="skip the several-streams-per-GPU variant")
    ap.add_argument("--no-stereo", action="store_true", help="skip the upstream-stereo (image pairs) variant")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, rank, world)
    else:
        run_ours(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
