#!/usr/bin/env python
"""tools/host_fuzz.py [n] [seed] -- random host-tensor calls (the drop-in path of lua/init.c) against the oracle, with
VOLSTN_HOST_SLICE_MB=1 so that small batches already split into several ramped slices on rotating streams.
Sampler (float / double, G = 4 / 3, strided views of grid / gradOut, onlyGrid / onlyVolume), grid generator, Cumsum,
affine sampler, cage warp (both arithmetic modes)."""
import os, sys
os.environ.setdefault("VOLSTN_HOST_SLICE_MB", "1")
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from alignet_b200 import nn as vnn, synth
from oracle import oracle
from tests._util import bits_equal, max_rel
from tests.test_gpu_rowwarp import cage_chain_oracle, random_affine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2024)
bad = 0
counts = {}
T = torch.from_numpy


def fail(kind, what, info):
    global bad
    bad += 1
    print("BAD", kind, what, info, flush=True)


for it in range(n):
    kind = str(rng.choice(["sampler", "sampler", "gridgen", "cumsum", "affine", "cage"]))
    counts[kind] = counts.get(kind, 0) + 1
    B = int(rng.integers(1, 48))
    oshape = (int(rng.integers(2, 20)), int(rng.integers(2, 20)), int(rng.integers(8, 48)))
    ishape = (int(rng.integers(1, 16)), int(rng.integers(1, 16)), int(rng.integers(1, 40)))
    if kind == "sampler":
        dt = np.float32 if rng.random() < 0.7 else np.float64
        G, C = int(rng.choice([4, 3])), int(rng.integers(1, 4))
        gk = str(rng.choice(["g1", "g2", "g3", "g4"]))
        mode = str(rng.choice(["both", "both", "onlyGrid", "onlyVolume"]))
        strided = rng.random() < 0.4
        info = dict(B=B, dt=np.dtype(dt).name, G=G, C=C, gk=gk, mode=mode, strided=strided, ishape=ishape, oshape=oshape)
        vol = synth.volume("v1", B, *ishape, C, rng, dt)
        if strided:
            big = synth.grid(gk, B, oshape[0] + 2, oshape[1], 2 * oshape[2] + 1, rng, dt, G=G, csz=4)
            gobig = synth.grad_output(B, oshape[0] + 2, oshape[1], 2 * oshape[2] + 1, C, rng, dt)
            ix = (slice(None), slice(1, 1 + oshape[0]), slice(None), slice(1, 1 + 2 * oshape[2], 2))
            grid, go = big[ix], gobig[ix]
        else:
            grid = synth.grid(gk, B, *oshape, rng, dt, G=G, csz=4)
            go = synth.grad_output(B, *oshape, C, rng, dt)
        gc, goc = np.ascontiguousarray(grid), np.ascontiguousarray(go)
        ref_out = oracle.sampler_forward(vol, gc, G)
        ref_gv, ref_gg = oracle.sampler_backward(vol, gc, goc, G)
        m = vnn.BilinearSamplerVolumetric(gridWidth=G, onlyGrid=mode == "onlyGrid", onlyVolume=mode == "onlyVolume")
        out = m.updateOutput([T(vol), T(grid)]).numpy()
        gv, gg = m.updateGradInput([T(vol), T(grid)], T(go))
        tol = 1e-4 if dt == np.float32 else 1e-12
        if not bits_equal(out, ref_out): fail(kind, "out", info)
        if mode != "onlyVolume" and not bits_equal(gg.numpy(), ref_gg): fail(kind, "gradGrid", info)
        if mode != "onlyGrid" and max_rel(gv.numpy(), ref_gv) > tol: fail(kind, "gradVol", (info, max_rel(gv.numpy(), ref_gv)))
        if mode == "onlyGrid" and float(gv.abs().max()) != 0.0: fail(kind, "gradVol not zero", info)
        if mode == "onlyVolume" and float(gg.abs().max()) != 0.0: fail(kind, "gradGrid not zero", info)
    elif kind == "gridgen":
        dt = np.float32 if rng.random() < 0.7 else np.float64
        info = dict(B=B, dt=np.dtype(dt).name, oshape=oshape)
        Tm = random_affine(B, rng).astype(dt)
        gen = vnn.VolumetricGridGenerator(*oshape)
        out = gen.updateOutput(T(Tm)).numpy()
        if np.max(np.abs(out - oracle.gridgen_forward(Tm, *oshape))) > 1e-6: fail(kind, "fwd", info)
        gg = rng.standard_normal(out.shape).astype(dt)
        gT = gen.updateGradInput(T(Tm), T(gg)).numpy()
        if max_rel(gT, oracle.gridgen_backward(gg)) > 2e-5: fail(kind, "bwd", (info, max_rel(gT, oracle.gridgen_backward(gg))))
    elif kind == "cumsum":
        N = int(rng.integers(1, 4))
        dims = [int(rng.integers(2, 14)) for _ in range(N)]
        x = rng.standard_normal([max(B, 2), N] + dims).astype(np.float32)
        cs = vnn.Cumsum()
        info = dict(B=B, dims=dims)
        if not bits_equal(cs.updateOutput(T(x)).numpy(), oracle.cumsum_forward(x)): fail(kind, "fwd", info)
        if not bits_equal(cs.updateGradInput(T(x), T(x)).numpy(), oracle.cumsum_backward(x)): fail(kind, "bwd", info)
    elif kind == "affine":
        C = int(rng.integers(1, 4))
        info = dict(B=B, C=C, ishape=ishape, oshape=oshape)
        vol = synth.volume("v1", B, *ishape, C, rng)
        Tm = random_affine(B, rng)
        go = synth.grad_output(B, *oshape, C, rng)
        grid = vnn.VolumetricGridGenerator(*oshape).updateOutput(T(Tm).cuda()).cpu().numpy()
        ref_out = oracle.sampler_forward(vol, grid)
        ref_gv, ref_gg = oracle.sampler_backward(vol, grid, go)
        m = vnn.AffineSamplerVolumetric(*oshape)
        out = m.updateOutput([T(vol), T(Tm)]).numpy()
        gv, gT = m.updateGradInput([T(vol), T(Tm)], T(go))
        if not bits_equal(out, ref_out): fail(kind, "out", info)
        if max_rel(gv.numpy(), ref_gv) > 1e-4: fail(kind, "gradVol", info)
        if max_rel(gT.numpy(), oracle.gridgen_backward(ref_gg)) > 2e-5: fail(kind, "gradT", info)
    else:
        cdims = tuple(int(rng.integers(1, 8)) for _ in range(3))
        info = dict(B=B, cdims=cdims, oshape=oshape)
        src = synth.volume("v2", B, *oshape, 1, rng)
        cage = np.stack([synth.alignet_cage(1, max(max(cdims), 2), rng)[0][:, :cdims[0], :cdims[1], :cdims[2]] for _ in range(B)]).astype(np.float32)
        go = synth.grad_output(B, *oshape, 1, rng)
        ref_out, ref_gsrc, ref_gcage = cage_chain_oracle(src, cage, oshape, go)
        m = vnn.CageWarpVolumetric()
        out = m.updateOutput([T(src), T(cage)]).numpy()
        gs, gc = m.updateGradInput([T(src), T(cage)], T(go))
        if not bits_equal(out, ref_out): fail(kind, "out", info)
        if max_rel(gs.numpy(), ref_gsrc) > 1e-4 or max_rel(gc.numpy(), ref_gcage) > 1e-4: fail(kind, "grads", info)
        mf = vnn.CageWarpVolumetric(fastArith=True)
        outf = mf.updateOutput([T(src), T(cage)]).numpy()
        if max_rel(outf, ref_out) > 3e-5: fail(kind, "fast out", (info, max_rel(outf, ref_out)))
print(f"{n} cases {counts}, {bad} failures")
