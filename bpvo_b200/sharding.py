"""Host-side logic of the point-sharded multi-GPU mode (no reference counterpart).

* `shard_range` -- the split `select_scan_kernel` applies on the device: contiguous scan-order blocks, multiples of 16.
* `radix_select_exchange` -- the exact-median protocol the ranks run per GN iteration, written against a generic
  `allreduce` callable (torch.distributed on CPU/gloo in the tests, NCCL on the device): three histogram exchanges over
  the float bit patterns of |r| (bits [30:20], [19:9], [8:0]) for the two middle ranks n/2-1 and n/2.
* `normal_equations_exchange` -- the 30 fp64 scalars (21 H + 6 G + sum w r^2 + good count + valid count).
* `bracket_select_exchange` -- the usual path of the peer-memory mode (the GN loop on the device, kernels_linearize.cuh
  `bracket_select_xrank`): ONE all-reduce of a 1024-bin linear histogram over a bracket around the previous median (+ the
  counts of valid residuals and of residuals below the bracket), then ONE all-gather of the handful of values in the two
  wanted bins.  Returns None when the bracket misses (the radix exchange above is the fallback).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, rank: int, size: int):
    """-> (first, n) of `rank`: groups of 16 points dealt out as contiguous blocks (floor split of the group count)."""
    groups = n_total // 16
    g0 = groups * rank // size
    g1 = groups * (rank + 1) // size
    return g0 * 16, (g1 - g0) * 16


def _bits(absr: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(absr, dtype=np.float32).view(np.uint32)


def radix_select_exchange(local_abs_residuals: np.ndarray, allreduce):
    """Exact (lo, hi) = order statistics n/2-1 and n/2 (lo == hi for odd n) of the union of all ranks' values.
    `allreduce(np.ndarray[int64]) -> np.ndarray` must return the element-wise sum over ranks."""
    bits = _bits(local_abs_residuals)
    h1 = allreduce(np.bincount(bits >> 20, minlength=2048).astype(np.int64))
    n = int(h1.sum())
    if n == 0:
        return 0, None, None
    t_hi = n // 2
    t_lo = t_hi - 1 if (n % 2 == 0 and n >= 2) else t_hi
    out = []
    c1 = np.cumsum(h1)
    sel1 = []
    for t in (t_lo, t_hi):
        b1 = int(np.searchsorted(c1, t, side="right"))
        sel1.append((b1, t - (int(c1[b1 - 1]) if b1 else 0)))
    h2 = allreduce(np.concatenate([np.bincount((bits[(bits >> 20) == b1] >> 9) & 2047, minlength=2048) for b1, _ in sel1]).astype(np.int64))
    sel2 = []
    for k, (b1, rem1) in enumerate(sel1):
        c2 = np.cumsum(h2[k * 2048:(k + 1) * 2048])
        b2 = int(np.searchsorted(c2, rem1, side="right"))
        sel2.append((b1, b2, rem1 - (int(c2[b2 - 1]) if b2 else 0)))
    h3 = allreduce(np.concatenate([np.bincount(bits[(bits >> 9) == ((b1 << 11) | b2)] & 511, minlength=512) for b1, b2, _ in sel2]).astype(np.int64))
    for k, (b1, b2, rem2) in enumerate(sel2):
        c3 = np.cumsum(h3[k * 512:(k + 1) * 512])
        b3 = int(np.searchsorted(c3, rem2, side="right"))
        out.append(np.array([(b1 << 20) | (b2 << 9) | b3], dtype=np.uint32).view(np.float32)[0])
    return n, out[0], out[1]


def median_from_pair(n: int, lo, hi) -> float:
    """median rule of bpvo/utils.h:224-252 for n >= 3"""
    return float(hi) if n % 2 else float(np.float32((np.float32(lo) + np.float32(hi)) / 2.0))


def normal_equations_exchange(J: np.ndarray, r: np.ndarray, w: np.ndarray, valid: np.ndarray, allreduce):
    """local fp64 sums of H (upper triangle, row-major), G, sum w r^2, valid count -> all-reduced vector of 30 scalars"""
    wv = w.astype(np.float64) * valid.astype(np.float64)
    Jd = J.astype(np.float64)
    H = (Jd * wv[:, None]).T @ Jd
    G = Jd.T @ (wv * r.astype(np.float64))
    vec = np.zeros(30)
    vec[:21] = H[np.triu_indices(6)]
    vec[21:27] = G
    vec[27] = float(np.sum(wv * r.astype(np.float64) ** 2))
    vec[29] = float(valid.sum())
    return allreduce(vec)


SEL_BINS = 1024


def _sel_bin(v: np.ndarray, lo: np.float32, inv_w: np.float32) -> np.ndarray:
    """kernels_linearize.cuh sel_bin(): min(kSelBins - 1, (int) ((v - lo) * inv_w)) in fp32"""
    t = (np.asarray(v, np.float32) - np.float32(lo)) * np.float32(inv_w)
    return np.minimum(SEL_BINS - 1, t.astype(np.int64))


def bracket_select_exchange(local_abs_residuals: np.ndarray, br_lo, br_hi, allreduce, allgather):
    """-> (n, lo, hi) like radix_select_exchange, or None on a bracket miss.  `allgather(np.ndarray[float32]) -> list`
    returns every rank's (variable-length) array."""
    a = np.ascontiguousarray(local_abs_residuals, dtype=np.float32)
    br_lo, br_hi = np.float32(br_lo), np.float32(br_hi)
    inv_w = np.float32(SEL_BINS) / (br_hi - br_lo) if br_hi > br_lo else np.float32(0.0)
    cand = a[(a >= br_lo) & (a <= br_hi)]
    msg = np.zeros(SEL_BINS + 3, np.int64)
    msg[:SEL_BINS] = np.bincount(_sel_bin(cand, br_lo, inv_w), minlength=SEL_BINS)
    msg[SEL_BINS:] = (a.size, int((a < br_lo).sum()), cand.size)
    g = allreduce(msg)
    n, below, ncand = int(g[SEL_BINS]), int(g[SEL_BINS + 1]), int(g[SEL_BINS + 2])
    if n < 3:
        return None
    t_hi = n // 2
    t_lo = t_hi - 1 if n % 2 == 0 else t_hi
    if below > t_lo or t_hi >= below + ncand:
        return None
    c = np.cumsum(g[:SEL_BINS])
    want = []
    for t in (t_lo - below, t_hi - below):
        b = int(np.searchsorted(c, t, side="right"))
        want.append((b, t - (int(c[b - 1]) if b else 0)))
    bins = _sel_bin(cand, br_lo, inv_w)
    mine = cand[(bins == want[0][0]) | (bins == want[1][0])]
    merged = np.concatenate([np.asarray(x, np.float32) for x in allgather(mine)])
    mb = _sel_bin(merged, br_lo, inv_w)
    out = [np.sort(merged[mb == b])[rem] for b, rem in want]
    return n, out[0], out[1]
