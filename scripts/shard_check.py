"""Point-sharded multi-GPU mode check, launched with torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/shard_check.py

Every rank feeds the SAME frames; rank r keeps its block of template points; H, G, sigma, the poses and the key-frame
decisions must equal those of an unsharded ctx on the same GPU.  Prints timing of the sharded vs unsharded linearize on the
dense 1080p workload (BASELINE.json configs[3])."""
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
    from bpvo_b200.engine import Context, FLAG_HOST_SOLVE
    from conftest import make_params, rel_err
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = [Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    report = {"world": world}

    # ---- parity: sharded vs unsharded on the same frames ------------------------------------------------------
    for name, sc, p in [("small-bitplanes", synth.scene_small(96, 128), make_params("bitplanes", 3, "tukey")),
                        ("kitti-bitplanes", synth.scene_kitti(), make_params("bitplanes", 4, "tukey")),
                        ("vga-intensity-l2", synth.scene_vga(), make_params("intensity", 3, "l2"))]:
        i0, d0 = sc.render(0); i1, d1 = sc.render(1)
        full = Context(sc.K, sc.baseline, (sc.rows, sc.cols), p, device_id=local, flags=FLAG_HOST_SOLVE)
        sh = Context(sc.K, sc.baseline, (sc.rows, sc.cols), p, device_id=local)
        uid2 = [Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid2, src=0)
        sh.comm_init(rank, world, uid2[0])
        fa, fb, sa, sb = full.frame(), full.frame(), sh.frame(), sh.frame()
        for a, b in ((fa, fb), (sa, sb)):
            a.setData(i0, d0); a.setTemplate(); b.setData(i1, d1)
        L = p.numPyramidLevels
        n_local = [sa.numPoints(l) for l in range(L)]
        n_full = [fa.numPoints(l) for l in range(L)]
        t = torch.tensor(n_local, device="cuda"); dist.all_reduce(t)
        assert t.tolist() == n_full, (t.tolist(), n_full)
        assert rel_err(sa.normalization(0), fa.normalization(0)) < 1e-6
        T = np.array(sc.relative_pose(0, 1), dtype=np.float32)
        for l in range(L - 1, -1, -1):
            g = full.linearize(fa, fb, l, T, True)
            s = sh.linearize(sa, sb, l, T, True)
            assert s["sigma"] == g["sigma"], (name, l, s["sigma"], g["sigma"])
            assert s["n_valid"] == g["n_valid"]
            assert rel_err(s["H"], g["H"]) < 2e-6 and np.abs(s["G"] - g["G"]).max() <= 2e-6 * np.abs(g["H"]).max() ** 0.5 * max(1.0, g["f_norm"]), (name, l)
            assert abs(s["f_norm"] - g["f_norm"]) <= 2e-6 * max(1.0, g["f_norm"])
            assert abs(sh.getFractionOfGoodPoints(p.goodPointThreshold) - full.getFractionOfGoodPoints(p.goodPointThreshold)) < 1e-6
        Tf, _, nf = full.estimatePose(fa, fb, np.eye(4, dtype=np.float32))
        Ts, _, ns = sh.estimatePose(sa, sb, np.eye(4, dtype=np.float32))
        assert rel_err(Ts, Tf) < 1e-4, (name, Ts, Tf)
        # all ranks hold the bit-identical pose
        tt = torch.from_numpy(Ts.copy()).cuda(); ref = tt.clone(); dist.broadcast(ref, src=0)
        assert torch.equal(tt, ref)
        report[name] = {"n_full": n_full, "n_local": n_local, "evals_full": nf, "evals_sharded": ns, "pose_rel_err": rel_err(Ts, Tf)}
        sh.comm_destroy()
        for o in (fa, fb, sa, sb):
            o.close()
        full.close(); sh.close()

    # ---- whole VisualOdometry streams stay in lock-step ----------------------------------------------------------
    sc = synth.scene_small(96, 128)
    p = make_params("bitplanes", 3, "tukey")
    vo = VisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p, device_id=local)
    vo.ctx.comm_init(rank, world, uid[0])
    for k in range(6):
        r = vo.addFrame(*sc.render(k))
        tt = torch.from_numpy(r.pose.copy()).cuda(); ref = tt.clone(); dist.broadcast(ref, src=0)
        assert torch.equal(tt, ref), f"frame {k}: ranks diverged"
    report["vo_stream"] = "lock-step over 6 frames"

    # ---- timing: dense 1080p (configs[3]) linearize, sharded vs one GPU -------------------------------------------
    sc = synth.scene_1080p()
    p = make_params("bitplanes", 5, "tukey", nonMaxSuppRadius=-1)
    i0, d0 = sc.render(0); i1, d1 = sc.render(1)
    T = np.eye(4, dtype=np.float32)
    times = {}
    for mode in ("single", "sharded"):
        ctx = Context(sc.K, sc.baseline, (sc.rows, sc.cols), p, device_id=local)
        if mode == "sharded":
            uid3 = [Context.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid3, src=0)
            ctx.comm_init(rank, world, uid3[0])
        a, b = ctx.frame(), ctx.frame()
        a.setData(i0, d0); a.setTemplate(); b.setData(i1, d1)
        n0 = a.numPoints(0)
        dist.barrier()
        ms = ctx.time_linearize(a, b, 0, T, iters=20, flush_l2=False)
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times[mode] = {"level0_points_local": n0, "ms_per_linearize": float(t[0])}
        if mode == "sharded":
            ctx.comm_destroy()
        a.close(); b.close(); ctx.close()
    report["dense_1080p_linearize"] = times
    report["dense_1080p_speedup"] = times["single"]["ms_per_linearize"] / times["sharded"]["ms_per_linearize"]
    if rank == 0:
        print("SHARD_CHECK_OK " + json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
