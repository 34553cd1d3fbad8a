"""Generate tests/golden/sampler_golden.npz from the REFERENCE ITSELF.

The reference ships no golden vectors, so these are made by running the
reference's own stn3dtorch/generic/BilinearSamplerVolumetric.c -- compiled unmodified by
oracle/Makefile (`make -C oracle ref`) behind oracle/ref_shim.c -- on small seeded inputs.
Only possible where /root/reference exists (the build container); the .npz is committed.

    python tests/golden/make_golden.py

Each case stores its inputs (vol, grid, gradOut) and the reference outputs (out, gradVol, gradGrid)
for the float and the double instantiation and both grid widths (G = 4: ...c:442-829,
G = 3: ...c:9-439).
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from alignet_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sampler_golden.npz")

# name: (B, (iD,iH,iW), C, (oD,oH,oW), grid kind, volume kind)
CASES = {
    "identity_8":      (2, (8, 8, 8), 1, (8, 8, 8), "g1", "v1"),
    "alignet_smooth":  (2, (9, 10, 11), 1, (9, 10, 11), "g2", "v2"),
    "uniform_c1":      (2, (5, 6, 7), 1, (4, 5, 9), "g3", "v1"),
    "uniform_c3":      (2, (5, 6, 7), 3, (4, 5, 9), "g3", "v1"),
    "uniform_c4":      (1, (6, 5, 4), 4, (7, 3, 5), "g3", "v1"),
    "uniform_c5":      (1, (4, 4, 4), 5, (3, 3, 3), "g3", "v1"),
    "extreme_c4":      (2, (6, 6, 6), 4, (5, 6, 7), "g4", "v1"),
    "extreme_c2":      (2, (3, 2, 2), 2, (6, 5, 4), "g4", "v2"),
    "affine_c1":       (2, (8, 7, 6), 1, (8, 7, 6), "g5", "v1"),
    "upsample_cage":   (2, (4, 4, 4), 4, (9, 9, 9), "g1", "v1"),
    "unit_extent":     (1, (1, 5, 1), 2, (3, 4, 5), "g3", "v1"),
    "id_64_floor":     (1, (2, 3, 64), 1, (2, 3, 64), "g1", "v1"),  # 23/64 identity coords floor 1 ulp low
}


def special_grid(dtype, G):
    """Hand-placed coordinates: exact +-1, just outside, huge, NaN/Inf (x86 cvttsd2si behaviour)."""
    vals = [-1.0, 1.0, -1.0000001, 1.0000001, 0.0, 0.9999999, 1e12, -1e12, 3e9, np.inf, -np.inf, np.nan, 2.5, -2.5]
    n = len(vals)
    g = np.zeros((1, 1, n, n, G), dtype)
    for i, a in enumerate(vals):
        for j, b in enumerate(vals):
            g[0, 0, i, j, 0] = a
            g[0, 0, i, j, 1] = b
            g[0, 0, i, j, 2] = vals[(i + j) % n]
    if G == 4:
        g[..., 3] = 1
    return g


def main():
    if not oracle.have_ref():
        oracle.build()
    assert oracle.have_ref(), "oracle/_ref is missing and /root/reference is not available"
    blob = {}
    seed = 1234
    for name, (B, ishape, C, oshape, gk, vk) in CASES.items():
        for G in (4, 3):
            for dt in (np.float32, np.float64):
                seed += 1
                rng = np.random.default_rng(seed)
                vol = synth.volume(vk, B, *ishape, C, rng, dt)
                grid = synth.grid(gk, B, *oshape, rng, dt, G=G, csz=4)
                go = synth.grad_output(B, *oshape, C, rng, dt)
                out = oracle.sampler_forward(vol, grid, G, "ref")
                gv, gg = oracle.sampler_backward(vol, grid, go, G, "ref")
                key = f"{name}|G{G}|{np.dtype(dt).name}"
                for k, v in (("vol", vol), ("grid", grid), ("gradOut", go), ("out", out), ("gradVol", gv), ("gradGrid", gg)):
                    blob[f"{key}|{k}"] = v
    # special coordinates; finite volume, so NaN/Inf come only from the grid
    for G in (4, 3):
        for dt in (np.float32, np.float64):
            rng = np.random.default_rng(99)
            vol = synth.volume("v1", 1, 3, 4, 5, 2, rng, dt)
            grid = special_grid(dt, G)
            go = synth.grad_output(1, *grid.shape[1:4], 2, rng, dt)
            out = oracle.sampler_forward(vol, grid, G, "ref")
            gv, gg = oracle.sampler_backward(vol, grid, go, G, "ref")
            key = f"special|G{G}|{np.dtype(dt).name}"
            for k, v in (("vol", vol), ("grid", grid), ("gradOut", go), ("out", out), ("gradVol", gv), ("gradGrid", gg)):
                blob[f"{key}|{k}"] = v
    np.savez_compressed(OUT, **blob)
    print(f"wrote {OUT}: {len(blob)} arrays, {os.path.getsize(OUT) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
