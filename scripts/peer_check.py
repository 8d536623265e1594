"""Peer-memory mode check (point-sharded estimate_pose with the GN loop ON THE DEVICE, exchanges over NVLink inside the
kernel), launched with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/peer_check.py

Every rank feeds the SAME frames and keeps its block of template points.  Checked: the sharded on-device solve agrees with
the unsharded on-device solve (pose <= 1e-4 relative, the north star's bound) and with the NCCL host-driven sharded solve;
all ranks hold bit-identical poses; whole VisualOdometry streams stay in lock-step.  Timed: one GN iteration of the dense
1080p workload (BASELINE.json configs[3]) on one GPU, NCCL-sharded (host loop) and peer-sharded (device loop)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from bpvo_b200 import VisualOdometry, synth
    from bpvo_b200.engine import Context
    from conftest import make_params, rel_err
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def uid():
        u = [Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(u, src=0)
        return u[0]

    def same_on_all_ranks(arr):
        t = torch.from_numpy(np.ascontiguousarray(arr).copy()).cuda(); ref = t.clone(); dist.broadcast(ref, src=0)
        return bool(torch.equal(t, ref))

    report = {"world": world}
    T0 = np.eye(4, dtype=np.float32)
    for name, sc, p in [("small-bitplanes-tukey", synth.scene_small(96, 128), make_params("bitplanes", 3, "tukey")),
                        ("small-intensity-huber", synth.scene_small(96, 128), make_params("intensity", 3, "huber")),
                        ("vga-intensity-l2", synth.scene_vga(), make_params("intensity", 3, "l2")),
                        ("kitti-bitplanes-tukey", synth.scene_kitti(), make_params("bitplanes", 4, "tukey"))]:
        i0, d0 = sc.render(0); i1, d1 = sc.render(1)
        full = Context(sc.K, sc.baseline, (sc.rows, sc.cols), p, device_id=local)
        nccl = Context(sc.K, sc.baseline, (sc.rows, sc.cols), p, device_id=local); nccl.comm_init(rank, world, uid())
        peer = Context(sc.K, sc.baseline, (sc.rows, sc.cols), p, device_id=local); peer.comm_init(rank, world, uid()); peer.peer_init_distributed(dist)
        peer.peer_set_min_points(0)            # shard EVERY level: the exchanges are what is under test here
        rep = Context(sc.K, sc.baseline, (sc.rows, sc.cols), p, device_id=local); rep.comm_init(rank, world, uid()); rep.peer_init_distributed(dist)
        rep.peer_set_min_points(1 << 30)       # ... and replicate every level: must equal the single-GPU run bit for bit
        out = {}
        for tag, ctx in (("full", full), ("nccl", nccl), ("peer", peer), ("rep", rep)):
            a, b = ctx.frame(), ctx.frame()
            a.setData(i0, d0); a.setTemplate(); b.setData(i1, d1)
            T, stats, n = ctx.estimatePose(a, b, T0)
            frac = ctx.getFractionOfGoodPoints(p.goodPointThreshold)
            out[tag] = (T, n, frac, [s.status for s in stats] if hasattr(stats[0], "status") else None)
            a.close(); b.close()
        Tf, Tn, Tp = out["full"][0], out["nccl"][0], out["peer"][0]
        assert rel_err(Tp, Tf) < 1e-4, (name, Tp, Tf)
        assert rel_err(Tp, Tn) < 1e-4, (name, Tp, Tn)
        assert same_on_all_ranks(Tp), f"{name}: ranks hold different poses"
        assert np.array_equal(out["rep"][0], Tf) and out["rep"][1] == out["full"][1], (name, "replicated levels must reproduce the single-GPU solve")
        assert abs(out["peer"][2] - out["full"][2]) < 5e-3, (name, out["peer"][2], out["full"][2])
        # host-driven paths on a peer-mode ctx with the DEFAULT shard_min_points (every level of these scenes is replicated):
        # linearize and the HOST_SOLVE loop must not sum the replicated levels over the ranks -- H, G, n_valid as on one GPU
        from bpvo_b200.engine import FLAG_HOST_SOLVE
        hs = Context(sc.K, sc.baseline, (sc.rows, sc.cols), p, device_id=local, flags=FLAG_HOST_SOLVE)
        hs.comm_init(rank, world, uid()); hs.peer_init_distributed(dist)
        a, b = hs.frame(), hs.frame(); a.setData(i0, d0); a.setTemplate(); b.setData(i1, d1)
        a1, b1 = full.frame(), full.frame(); a1.setData(i0, d0); a1.setTemplate(); b1.setData(i1, d1)
        if a.numPoints(0) == a1.numPoints(0):      # replicated (under 131072 points)
            lh, lf = hs.linearize(a, b, 0, T0, True), full.linearize(a1, b1, 0, T0, True)
            assert lh["n_valid"] == lf["n_valid"] and rel_err(lh["H"], lf["H"]) < 1e-6 and abs(lh["f_norm"] - lf["f_norm"]) <= 1e-6 * max(1.0, lf["f_norm"]), (name, lh, lf)
            Th, sh_, nh = hs.estimatePose(a, b, T0)
            assert rel_err(Th, Tf) < 1e-4, (name, "HOST_SOLVE on replicated levels", Th, Tf)
            assert hs.getFractionOfGoodPoints(p.goodPointThreshold) <= 1.0
        a.close(); b.close(); a1.close(); b1.close(); hs.comm_destroy(); hs.close()
        report[name] = {"evals_full": out["full"][1], "evals_nccl": out["nccl"][1], "evals_peer": out["peer"][1],
                        "pose_rel_err_vs_full": rel_err(Tp, Tf), "pose_rel_err_vs_nccl": rel_err(Tp, Tn)}
        nccl.comm_destroy(); peer.comm_destroy(); rep.comm_destroy()
        full.close(); nccl.close(); peer.close(); rep.close()

    # ---- whole VisualOdometry streams: lock-step, and the same trajectory as one GPU -------------------------------
    sc = synth.scene_small(96, 128)
    p = make_params("bitplanes", 3, "tukey")
    vo1 = VisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p, device_id=local)
    vo = VisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p, device_id=local)
    vo.ctx.comm_init(rank, world, uid()); vo.ctx.peer_init_distributed(dist); vo.ctx.peer_set_min_points(0)
    worst = 0.0
    for k in range(8):
        img, d = sc.render(k)
        r1 = vo1.addFrame(img, d)
        r = vo.addFrame(img, d)
        assert same_on_all_ranks(r.pose), f"frame {k}: ranks diverged"
        assert r.isKeyFrame == r1.isKeyFrame, f"frame {k}: key-frame decision differs from the single-GPU run"
        worst = max(worst, rel_err(r.pose, r1.pose))
    assert worst < 1e-4, worst
    report["vo_stream"] = {"frames": 8, "worst_pose_rel_err_vs_single_gpu": worst}
    vo.ctx.comm_destroy()

    # ---- sequence-number wrap-around of the cross-rank mailboxes: collective reset (test hook BPVO_B200_SEQ_INIT) ----------
    os.environ["BPVO_B200_SEQ_INIT"] = "0xfffff800"
    vo2 = VisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p, device_id=local)
    del os.environ["BPVO_B200_SEQ_INIT"]
    vo2.ctx.comm_init(rank, world, uid()); vo2.ctx.peer_init_distributed(dist); vo2.ctx.peer_set_min_points(0)
    vo3 = VisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p, device_id=local)
    vo3.ctx.comm_init(rank, world, uid()); vo3.ctx.peer_init_distributed(dist); vo3.ctx.peer_set_min_points(0)
    for k in range(4):
        img, d = sc.render(k)
        ra, rb = vo2.addFrame(img, d), vo3.addFrame(img, d)
        assert np.array_equal(ra.pose, rb.pose), f"frame {k}: the mailbox reset changed the result"
    report["sequence_reset"] = "ok"
    vo2.ctx.comm_destroy(); vo3.ctx.comm_destroy()

    # ---- timing: dense 1080p (configs[3]), one GN iteration ------------------------------------------------------------
    sc = synth.scene_1080p()
    p = make_params("bitplanes", 5, "tukey", nonMaxSuppRadius=-1)
    i0, d0 = sc.render(0); i1, d1 = sc.render(1)
    times = {}
    for mode in ("single", "nccl_host_loop", "peer_device_loop_shard_all_levels", "peer_device_loop"):
        ctx = Context(sc.K, sc.baseline, (sc.rows, sc.cols), p, device_id=local)
        if mode != "single":
            ctx.comm_init(rank, world, uid())
        if mode.startswith("peer"):
            ctx.peer_init_distributed(dist)
            if "shard_all" in mode:
                ctx.peer_set_min_points(0)
        a, b = ctx.frame(), ctx.frame()
        a.setData(i0, d0); a.setTemplate(); b.setData(i1, d1)
        ctx.estimatePose(a, b, T0)                       # warm-up
        ctx.set_profiling(True); ctx.reset_counters()
        dist.barrier()
        evals = 0
        for _ in range(3):
            T, _, n = ctx.estimatePose(a, b, T0)
            evals += n
        ms = ctx.counters()["ms_linearize"]
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times[mode] = {"points_local_per_level": [a.numPoints(l) for l in range(p.numPyramidLevels)], "gn_iterations": evals, "us_per_gn_iteration": 1e3 * float(t[0]) / max(evals, 1),
                       "pose_same_on_all_ranks": same_on_all_ranks(T)}
        if mode != "single":
            ctx.comm_destroy()
        a.close(); b.close(); ctx.close()
    report["dense_1080p"] = times
    if rank == 0:
        print("PEER_CHECK_OK " + json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
