"""Deterministic synthetic stereo-VO scenes (image + disparity + ground-truth motion).

The reference ships no data (its loaders read KITTI / Tsukuba from disk, utils/dataset.cc), so the
parity tests and the bench use this generator: a slanted, textured plane seen by a pin-hole stereo
rig that moves by a constant SE(3) step per frame.  Everything is IEEE basic arithmetic
(+ - * / sqrt floor) on float64 -- no sin/cos/exp -- so the same seed yields bit-identical u8 images
and f32 disparities on any host (the golden fixtures under tests/golden/ rely on that).

Texture: multi-octave value noise on lattices hashed with splitmix64, smoothstep-interpolated.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def _hash01(ix: np.ndarray, iy: np.ndarray, seed: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        h = (ix.astype(np.int64).view(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) ^ \
            (iy.astype(np.int64).view(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)) ^ np.uint64(seed & 0xFFFFFFFFFFFFFFFF)
        h = splitmix64(h)
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / float(1 << 53))


def value_noise(u: np.ndarray, v: np.ndarray, spacing: float, seed: int) -> np.ndarray:
    x = u / spacing
    y = v / spacing
    ix = np.floor(x)
    iy = np.floor(y)
    fx = x - ix
    fy = y - iy
    sx = fx * fx * (3.0 - 2.0 * fx)
    sy = fy * fy * (3.0 - 2.0 * fy)
    ix = ix.astype(np.int64)
    iy = iy.astype(np.int64)
    a = _hash01(ix, iy, seed)
    b = _hash01(ix + 1, iy, seed)
    c = _hash01(ix, iy + 1, seed)
    d = _hash01(ix + 1, iy + 1, seed)
    top = a + (b - a) * sx
    bot = c + (d - c) * sx
    return top + (bot - top) * sy


def cayley_se3(xi) -> np.ndarray:
    """SE(3) matrix from a small twist (w, v) using the Cayley map for the rotation (no sin/cos)."""
    w = [float(xi[0]) * 0.5, float(xi[1]) * 0.5, float(xi[2]) * 0.5]
    S = [[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]]
    A = [[(1.0 if i == j else 0.0) - S[i][j] for j in range(3)] for i in range(3)]
    B = [[(1.0 if i == j else 0.0) + S[i][j] for j in range(3)] for i in range(3)]
    # explicit adjugate inverse and products on python floats: plain IEEE doubles, no BLAS/FMA
    c = [[A[(i + 1) % 3][(j + 1) % 3] * A[(i + 2) % 3][(j + 2) % 3] - A[(i + 1) % 3][(j + 2) % 3] * A[(i + 2) % 3][(j + 1) % 3]
          for j in range(3)] for i in range(3)]
    det = A[0][0] * c[0][0] + A[0][1] * c[0][1] + A[0][2] * c[0][2]
    Ainv = [[c[j][i] / det for j in range(3)] for i in range(3)]
    R = _mm(Ainv, B)
    T = np.eye(4)
    T[:3, :3] = np.array(R)
    T[:3, 3] = np.asarray(xi[3:], dtype=np.float64)
    return T


def _mm(a, b):
    n, m, k = len(a), len(b[0]), len(b)
    out = [[0.0] * m for _ in range(n)]
    for i in range(n):
        for j in range(m):
            s = 0.0
            for q in range(k):
                s = s + float(a[i][q]) * float(b[q][j])
            out[i][j] = s
    return out


@dataclass
class Scene:
    rows: int
    cols: int
    fx: float
    cx: float
    cy: float
    baseline: float
    seed: int = 0xB200
    xi: tuple = (0.002, -0.003, 0.001, 0.02, -0.01, 0.04)   # per-frame camera motion (rad, m)
    plane_n: tuple = (0.10, -0.15, 1.0)
    plane_depth: float = 10.0          # depth of the plane along the first optical axis (m)
    hole_fraction: float = 0.0         # fraction of disparity holes (d = 0)
    octaves: tuple = (0.02, 0.04, 0.08, 0.16, 0.32, 0.64, 1.28)
    _M: np.ndarray = field(default=None, repr=False)

    @property
    def K(self) -> np.ndarray:
        return np.array([[self.fx, 0, self.cx], [0, self.fx, self.cy], [0, 0, 1]], dtype=np.float32)

    def cam_to_world(self, k: int) -> np.ndarray:
        if self._M is None:
            self._M = cayley_se3(self.xi)
        W = np.eye(4).tolist()
        M = self._M.tolist()
        for _ in range(k):
            W = _mm(W, M)
        return np.array(W)

    def relative_pose(self, k0: int, k1: int) -> np.ndarray:
        """Ground-truth T with X_cam(k1) = T @ X_cam(k0)."""
        return np.linalg.inv(self.cam_to_world(k1)) @ self.cam_to_world(k0)

    def render(self, k: int):
        """-> (image u8 [rows, cols], disparity f32 [rows, cols]) of frame k."""
        W = self.cam_to_world(k)
        return self._render_from(W[:3, :3], W[:3, 3], k)

    def render_right(self, k: int):
        """-> image u8 of the RIGHT camera of the rig at frame k (same orientation, centre one baseline along the camera's
        x axis): with render(k)[0] the rectified pair a stereo matcher turns into render(k)[1] (upstream stereo)."""
        W = self.cam_to_world(k)
        R, t = W[:3, :3], W[:3, 3]
        tr = np.array([t[0] + R[0, 0] * self.baseline, t[1] + R[1, 0] * self.baseline, t[2] + R[2, 0] * self.baseline])
        return self._render_from(R, tr, k)[0]

    def _render_from(self, R, t, k: int):
        n = np.asarray(self.plane_n, dtype=np.float64)
        n = n / np.sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2])
        d0 = n[2] * self.plane_depth
        a = np.array([1.0, 0.0, 0.0])
        e1 = a - n * n[0]
        e1 = e1 / np.sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2])
        e2 = np.array([n[1] * e1[2] - n[2] * e1[1], n[2] * e1[0] - n[0] * e1[2], n[0] * e1[1] - n[1] * e1[0]])
        xs = (np.arange(self.cols, dtype=np.float64) - self.cx) / self.fx
        ys = (np.arange(self.rows, dtype=np.float64) - self.cy) / self.fx
        dx, dy = np.meshgrid(xs, ys)
        # world ray direction R @ (dx, dy, 1)
        wx = R[0, 0] * dx + R[0, 1] * dy + R[0, 2]
        wy = R[1, 0] * dx + R[1, 1] * dy + R[1, 2]
        wz = R[2, 0] * dx + R[2, 1] * dy + R[2, 2]
        s = (d0 - (n[0] * t[0] + n[1] * t[1] + n[2] * t[2])) / (n[0] * wx + n[1] * wy + n[2] * wz)   # = depth along the optical axis
        Xw = t[0] + s * wx
        Yw = t[1] + s * wy
        Zw = t[2] + s * wz
        u = e1[0] * Xw + e1[1] * Yw + e1[2] * Zw
        v = e2[0] * Xw + e2[1] * Yw + e2[2] * Zw
        acc = np.zeros_like(u)
        wsum = 0.0
        for i, sp in enumerate(self.octaves):
            amp = 1.0 + 0.25 * i
            acc = acc + amp * value_noise(u, v, sp, self.seed + 7919 * (i + 1))
            wsum += amp
        tex = acc / wsum
        # stretch contrast around the mean (value-noise sums concentrate near 0.5)
        tex = 0.5 + (tex - 0.5) * 2.6
        img = np.clip(np.rint(16.0 + 224.0 * tex), 0, 255).astype(np.uint8)
        disp = (self.fx * self.baseline / s).astype(np.float32)
        if self.hole_fraction > 0.0:
            yy, xx = np.meshgrid(np.arange(self.rows, dtype=np.int64), np.arange(self.cols, dtype=np.int64), indexing="ij")
            h = _hash01(xx, yy, self.seed + 1 + 104729 * k)
            disp = np.where(h < self.hole_fraction, np.float32(0.0), disp).astype(np.float32)
        return np.ascontiguousarray(img), np.ascontiguousarray(disp)


# the BASELINE.json configurations
def scene_vga(seed: int = 0xB200, **kw) -> Scene:
    return Scene(rows=480, cols=640, fx=615.0, cx=320.0, cy=240.0, baseline=0.1, seed=seed, plane_depth=4.0,
                 xi=(0.002, -0.003, 0.001, 0.008, -0.004, 0.016), octaves=(0.008, 0.016, 0.032, 0.064, 0.128, 0.256, 0.512), **kw)


def scene_kitti(seed: int = 0xB200, **kw) -> Scene:
    return Scene(rows=376, cols=1241, fx=718.856, cx=607.1928, cy=185.2157, baseline=0.5372, seed=seed, **kw)


def scene_1080p(seed: int = 0xB200, **kw) -> Scene:
    return Scene(rows=1080, cols=1920, fx=1400.0, cx=960.0, cy=540.0, baseline=0.12, seed=seed, plane_depth=6.0,
                 xi=(0.002, -0.003, 0.001, 0.012, -0.006, 0.024), octaves=(0.005, 0.01, 0.02, 0.04, 0.08, 0.16, 0.32, 0.64), **kw)


def scene_small(rows: int = 96, cols: int = 128, seed: int = 7, **kw) -> Scene:
    """Tiny scene for CPU-side unit tests."""
    return Scene(rows=rows, cols=cols, fx=120.0, cx=cols / 2.0, cy=rows / 2.0, baseline=0.1, seed=seed, plane_depth=3.0,
                 xi=(0.002, -0.003, 0.001, 0.006, -0.003, 0.012), octaves=(0.03, 0.06, 0.12, 0.24, 0.48), **kw)
