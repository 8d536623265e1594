"""world_size-2 gloo test (CPU) of the host-side logic of the point-sharded multi-GPU mode: the shard split, the
exact-median histogram exchange and the 30-scalar normal-equation exchange, against the oracle's single-process
result on the same synthetic pair."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, make_params


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from bpvo_b200 import sharding, synth
    from oracle import pyoracle as po
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def allreduce(a):
            t = torch.from_numpy(np.ascontiguousarray(a).copy())
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return t.numpy()

        sc = synth.scene_small(96, 128)
        p = make_params("bitplanes", 2, "tukey")
        a = po.Frame(sc.K, sc.baseline, sc.rows, sc.cols, p, use_rcp=0)
        b = po.Frame(sc.K, sc.baseline, sc.rows, sc.cols, p, use_rcp=0)
        i0, d0 = sc.render(0); i1, d1 = sc.render(1)
        a.set_data(i0, d0); a.set_template(); b.set_data(i1, d1)
        T = np.eye(4, dtype=np.float32); T[0, 3] = 0.05
        full = po.Estimator(p).linearize(a, b, 0, T)            # every rank computes the full problem as the checker
        N, Cch = a.num_points(0), 8
        first, n = sharding.shard_range(N, rank, world)
        sl = slice(first, first + n)
        r = full["residuals"].reshape(Cch, N)[:, sl]
        v = full["valid"].reshape(Cch, N)[:, sl]
        w = full["weights"].reshape(Cch, N)[:, sl]
        J = a.jacobians(0)[:, sl, :]
        # 1. exact global median from three histogram exchanges
        absr = np.abs(r[v != 0])
        n_valid, lo, hi = sharding.radix_select_exchange(absr, allreduce)
        all_abs = np.sort(np.abs(full["residuals"][full["valid"] != 0]))
        assert n_valid == all_abs.size
        assert lo == all_abs[n_valid // 2 - 1 if n_valid % 2 == 0 else n_valid // 2] and hi == all_abs[n_valid // 2]
        sigma = np.float32(np.float32(1.4826) * (np.float32(1.0) + np.float32(5.0) / np.float32(n_valid - 6))) * np.float32(sharding.median_from_pair(n_valid, lo, hi))
        assert sigma == np.float32(full["sigma"])
        # 1b. the peer-memory mode's usual path: bracket histogram all-reduce + short-list all-gather (exact when it hits)
        def allgather(x):
            out = [None] * world
            dist.all_gather_object(out, np.asarray(x, np.float32))
            return out
        med = np.float32(0.5) * (np.float32(lo) + np.float32(hi))
        hit = sharding.bracket_select_exchange(absr, med * np.float32(0.97), med * np.float32(1.03), allreduce, allgather)
        assert hit is not None and hit[0] == n_valid and hit[1] == lo and hit[2] == hi, hit
        miss = sharding.bracket_select_exchange(absr, med * np.float32(1.5), med * np.float32(1.6), allreduce, allgather)
        assert miss is None
        # 2. the 30 normal-equation scalars
        vec = sharding.normal_equations_exchange(J.reshape(-1, 6), r.ravel(), w.ravel(), v.ravel(), allreduce)
        H = np.zeros((6, 6)); H[np.triu_indices(6)] = vec[:21]; H = H + np.triu(H, 1).T
        assert np.abs(H - full["H"]).max() <= 2e-4 * np.abs(full["H"]).max()
        assert np.abs(vec[21:27] - full["G"]).max() <= 2e-3 * np.abs(full["G"]).max() + 1e-3
        assert abs(np.sqrt(vec[27]) - full["f_norm"]) <= 1e-3 * full["f_norm"]
        assert int(vec[29]) == int(full["valid"].sum())
        q.put((rank, "ok", first, n, N))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "fail: " + traceback.format_exc(), 0, 0, 0))
    finally:
        dist.destroy_process_group()


def test_shard_ranges_partition_the_points():
    from bpvo_b200 import sharding
    for n_total in [0, 16, 48, 29424, 401040, 2052640]:
        for size in [1, 2, 3, 4, 8]:
            pos = 0
            for rank in range(size):
                first, n = sharding.shard_range(n_total, rank, size)
                assert first == pos and n % 16 == 0
                pos += n
            assert pos == n_total // 16 * 16


def test_two_rank_exchange_matches_single_process(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    res.sort()
    for rank, status, first, n, N in res:
        assert status == "ok", status
    assert res[0][2] == 0 and res[0][3] + res[1][3] == res[0][4]
