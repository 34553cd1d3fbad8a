"""Generate tests/golden/cage_warp_golden.npz from the REFERENCE ITSELF.

The fused modules stand for chains of reference modules.  Their golden
vectors are made by running that chain with the reference's own stn3dtorch/generic/BilinearSamplerVolumetric.c
(compiled unmodified, oracle/_ref, `impl="ref"`) for every sampler call:

    fieldI = the generator's base grid (VolumetricGridGenerator.lua:12-23)
    field  = ref.updateOutput(cage volume B x cD x cH x cW x 4, fieldI)          -- the cage up-sample
    out    = ref.updateOutput(source, field)                                     -- the source warp
    gradSource, gradField = ref.updateGradInput(source, field, gradOut)
    gradCage = ref.updateGradInput(cage volume, fieldI, gradField)[0][..., :3]

and, for the one-kernel training step, gradOut = SmoothL1Criterion gradient of (out, target) (numpy restatement of
THNN, oracle.smooth_l1_backward) with the loss and the IoU counts.  Only possible where /root/reference exists.

    python tests/golden/make_cage_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from alignet_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cage_warp_golden.npz")

# name: (B, cage dims, source dims, output dims, C, cage kind)
CASES = {
    "alignet_8":   (2, (8, 8, 8), (16, 16, 16), (16, 16, 16), 1, "alignet"),
    "alignet_4":   (2, (4, 4, 4), (12, 10, 14), (12, 10, 14), 1, "alignet"),
    "identity_6":  (1, (6, 6, 6), (12, 12, 12), (12, 12, 12), 1, "identity"),
    "folded":      (2, (3, 5, 2), (7, 6, 9), (9, 6, 37), 2, "random"),
    "wide_row":    (1, (4, 3, 12), (5, 6, 40), (4, 5, 70), 1, "alignet"),
}


def chain(src, cage, oshape, go, impl):
    B = src.shape[0]
    base = oracle.gridgen_base(*oshape)
    field_i = np.ascontiguousarray(np.broadcast_to(base, (B,) + base.shape))
    cage_vol = np.ones(cage.shape[:1] + cage.shape[2:] + (4,), np.float32)
    cage_vol[..., :3] = np.moveaxis(cage, 1, -1)
    field = oracle.sampler_forward(cage_vol, field_i, 4, impl)
    out = oracle.sampler_forward(src, field, 4, impl)
    if go is None:
        return field, out
    gsrc, gfield = oracle.sampler_backward(src, field, go, 4, impl)
    gcv, _ = oracle.sampler_backward(cage_vol, field_i, gfield, 4, impl)
    return field, out, gsrc, np.ascontiguousarray(np.moveaxis(gcv[..., :3], -1, 1))


def make_cage(kind, B, cdims, rng):
    if kind == "identity":
        ax = [np.linspace(-1, 1, n) for n in cdims]
        c = np.empty((B, 3) + cdims, np.float32)
        c[:, 0], c[:, 1], c[:, 2] = ax[0][:, None, None], ax[1][None, :, None], ax[2][None, None, :]
        return c
    if kind == "random":
        return rng.uniform(-1.5, 1.5, (B, 3) + cdims).astype(np.float32)
    m = max(cdims)
    return np.stack([synth.alignet_cage(1, m, rng)[0][:, :cdims[0], :cdims[1], :cdims[2]] for _ in range(B)]).astype(np.float32)


def main():
    if not oracle.have_ref():
        oracle.build()
    assert oracle.have_ref(), "oracle/_ref is missing and /root/reference is not available"
    blob = {}
    seed = 4321
    for name, (B, cdims, sdims, oshape, C, kind) in CASES.items():
        seed += 1
        rng = np.random.default_rng(seed)
        src = synth.volume("v2", B, *sdims, C, rng)
        src[0] *= 2.5  # |out - target| > 1 somewhere: the linear branch of the criterion
        cage = make_cage(kind, B, cdims, rng)
        go = synth.grad_output(B, *oshape, C, rng)
        target = synth.volume("v2", B, *oshape, C, rng)
        field, out, gsrc, gcage = chain(src, cage, oshape, go, "ref")
        blob.update({f"{name}|source": src, f"{name}|cage": cage, f"{name}|gradOut": go, f"{name}|target": target,
                     f"{name}|field": field, f"{name}|out": out, f"{name}|gradSource": gsrc, f"{name}|gradCage": gcage})
        if C == 1:  # the training step (single-channel volumes)
            gl = oracle.smooth_l1_backward(out, target)
            _, _, gsrc_s, gcage_s = chain(src, cage, oshape, gl, "ref")
            blob.update({f"{name}|step_loss": np.array(oracle.smooth_l1_forward(out, target), np.float64),
                         f"{name}|step_iou": oracle.iou_counts(out, target),
                         f"{name}|step_gradSource": gsrc_s, f"{name}|step_gradCage": gcage_s})
    np.savez_compressed(OUT, **blob)
    print(f"wrote {OUT}: {len(blob)} arrays, {os.path.getsize(OUT) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
