"""Synthetic inputs for tests and benchmarks: numpy, seeded, no dataset.

Volumes   v1  N(0,1)                       -- exercises the arithmetic
          v2  binary occupancy randn>=0.5  -- as stn3dtorch/test.lua:83
Grids     g1  identity field (grid generator with T = I) = ALIGNet's random-init warp
              (models/alignet_2d.lua:118-119)
          g2  ALIGNet-like: cage deltas |del*(1+0.25 N)| -> Cumsum -> +beta -> up-sampled to the
              output resolution (monotone, smooth; models/alignet_2d.lua:77-112)
          g3  U(-1,1)^3                    -- worst-case locality
          g4  extreme: U(-3,3) mixed with exact +-1 and a constant point (out of bounds + scatter
              contention; BASELINE.json config 5)
          g5  affine: random rotations through the grid generator (stn3dtorch/test.lua:27-81)
This module only manufactures inputs; it is not part of the computation being measured.
"""
from __future__ import annotations

import numpy as np


def base_axis(n: int) -> np.ndarray:
    """-1 + k/(n-1)*2 in double (VolumetricGridGenerator.lua:17-19)."""
    return -1.0 + np.arange(n, dtype=np.float64) / float(n - 1) * 2.0


def volume(kind: str, B: int, D: int, H: int, W: int, C: int, rng, dtype=np.float32) -> np.ndarray:
    x = rng.standard_normal((B, D, H, W, C), dtype=np.float32)
    if kind == "v1":
        return x.astype(dtype)
    if kind == "v2":
        return (x >= 0.5).astype(dtype)
    raise ValueError(kind)


def _upsample_axis(a: np.ndarray, axis: int, n_out: int) -> np.ndarray:
    n_in = a.shape[axis]
    pos = np.linspace(0.0, n_in - 1.0, n_out)
    i0 = np.minimum(np.floor(pos).astype(np.int64), n_in - 2) if n_in > 1 else np.zeros(n_out, np.int64)
    w = (pos - i0).astype(a.dtype)
    lo = np.take(a, i0, axis=axis)
    hi = np.take(a, np.minimum(i0 + 1, n_in - 1), axis=axis)
    shape = [1] * a.ndim
    shape[axis] = n_out
    w = w.reshape(shape)
    return lo * (1 - w) + hi * w


def alignet_cage(B: int, csz: int, rng, jitter: float = 0.25) -> np.ndarray:
    """B x 3 x csz^3 absolute cage coordinates: |del(1+jitter N)| -> cumsum along own axis -> + beta."""
    delta = 2.0 / (csz - 1)
    beta = -1.0 - delta
    d = np.abs(delta * (1.0 + jitter * rng.standard_normal((B, 3, csz, csz, csz))))
    cage = np.empty_like(d)
    for i in range(3):
        cage[:, i] = np.cumsum(d[:, i], axis=1 + i)
    return cage + beta


def grid(kind: str, B: int, oD: int, oH: int, oW: int, rng, dtype=np.float32, G: int = 4, csz: int = 8) -> np.ndarray:
    g = np.empty((B, oD, oH, oW, G), np.float64)
    if G == 4:
        g[..., 3] = 1.0
    if kind == "g1":
        g[..., 0] = base_axis(oD)[None, :, None, None]
        g[..., 1] = base_axis(oH)[None, None, :, None]
        g[..., 2] = base_axis(oW)[None, None, None, :]
    elif kind == "g2":
        cage = alignet_cage(B, csz, rng)  # B x 3 x csz^3
        up = cage
        for ax, n in zip((2, 3, 4), (oD, oH, oW)):
            up = _upsample_axis(up, ax, n)
        g[..., :3] = np.moveaxis(up, 1, -1)
    elif kind == "g3":
        g[..., :3] = rng.uniform(-1.0, 1.0, (B, oD, oH, oW, 3))
    elif kind == "g4":
        g[..., :3] = rng.uniform(-3.0, 3.0, (B, oD, oH, oW, 3))
        sel = rng.uniform(size=(B, oD, oH, oW))
        pm = np.where(rng.uniform(size=(B, oD, oH, oW, 3)) < 0.5, -1.0, 1.0)
        m = sel < 0.05
        g[..., :3][m] = pm[m]                      # exact corners of the volume
        m = (sel >= 0.05) & (sel < 0.06)
        g[..., :3][m] = np.array([0.25, -0.5, 0.75])  # one hot spot -> atomics contention
        m = (sel >= 0.06) & (sel < 0.30)
        g[..., :3][m] = rng.uniform(-1.0, 1.0, (int(m.sum()), 3))
    elif kind == "g5":
        q = rng.standard_normal((B, 3, 3))
        R, _ = np.linalg.qr(q)
        base = np.stack(np.meshgrid(base_axis(oD), base_axis(oH), base_axis(oW), indexing="ij"), -1)
        g[..., :3] = np.einsum("dhwk,bik->bdhwi", base, R)
    else:
        raise ValueError(kind)
    return g.astype(dtype)


def grad_output(B: int, oD: int, oH: int, oW: int, C: int, rng, dtype=np.float32) -> np.ndarray:
    return rng.standard_normal((B, oD, oH, oW, C), dtype=np.float32).astype(dtype)
