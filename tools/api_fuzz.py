#!/usr/bin/env python
"""tools/api_fuzz.py [n] [seed] -- random shapes, layouts and fields through the bit-exact entry points against the oracle.
Developer stress run (the committed tests cover fixed shapes).  Per case one of:
  sampler   BilinearSamplerVolumetric, G in {4, 3}, C in 1..6, layouts BDHWC / BCDHW, algos auto / row / direct, fields g1..g5
  cage      CageWarpVolumetric + the one-kernel step (bit-exact mode) against the oracle chain
  affine    AffineSamplerVolumetric against grid generator + sampler
  cumsum    nn.Cumsum forward / backward
Contract checked: out, gradGrid bit-exact; gradVol / gradCage <= 1e-4 max; gradT <= 1e-5; loss 1e-9; IoU exact."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from alignet_b200 import nn as vnn, synth, _abi
from oracle import oracle
from tests._util import bits_equal, max_rel
from tests.test_gpu_rowwarp import cage_chain_oracle, random_affine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 4321)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
host = lambda t: t.detach().contiguous().cpu().numpy()
bad = 0
counts = {}


def fail(kind, what, info):
    global bad
    bad += 1
    print("BAD", kind, what, info, flush=True)


for it in range(n):
    kind = rng.choice(["sampler", "sampler", "cage", "affine", "cumsum"])
    counts[kind] = counts.get(kind, 0) + 1
    B = int(rng.integers(1, 5))
    oshape = (int(rng.integers(1, 14)), int(rng.integers(1, 14)), int(rng.integers(1, 72)))
    ishape = (int(rng.integers(1, 14)), int(rng.integers(1, 14)), int(rng.integers(1, 40)))
    if kind == "sampler":
        G, C = int(rng.choice([4, 3])), int(rng.integers(1, 7))
        gk = str(rng.choice(["g1", "g2", "g3", "g4", "g5"]))
        algo = int(rng.choice([0, _abi.ALGO_ROW, _abi.ALGO_DIRECT]))
        layout = str(rng.choice(["BDHWC", "BCDHW"]))
        info = dict(B=B, G=G, C=C, gk=gk, algo=algo, layout=layout, ishape=ishape, oshape=oshape)
        vol = synth.volume("v1", B, *ishape, C, rng)
        if min(oshape) < 2 and gk in ("g2", "g5"):
            gk = "g3"
        grid = synth.grid(gk, B, *oshape, rng, G=G, csz=4)
        go = synth.grad_output(B, *oshape, C, rng)
        ref_out = oracle.sampler_forward(vol, grid, G)
        ref_gv, ref_gg = oracle.sampler_backward(vol, grid, go, G)
        m = vnn.BilinearSamplerVolumetric(gridWidth=G, algo=algo, layout=layout)
        if layout == "BCDHW":
            tv, tg, tgo = (dev(np.moveaxis(x, -1, 1)) for x in (vol, grid, go))
            back = lambda t: np.ascontiguousarray(np.moveaxis(host(t), 1, -1))
        else:
            tv, tg, tgo = dev(vol), dev(grid), dev(go)
            back = host
        out = back(m.updateOutput([tv, tg]))
        gv, gg = m.updateGradInput([tv, tg], tgo)
        if not bits_equal(out, ref_out): fail(kind, "out", info)
        if not bits_equal(back(gg), ref_gg): fail(kind, "gradGrid", info)
        if max_rel(back(gv), ref_gv) > 1e-4: fail(kind, "gradVol", (info, max_rel(back(gv), ref_gv)))
    elif kind == "cage":
        oshape = tuple(max(2, s) for s in oshape)
        C = int(rng.choice([1, 1, 2]))
        cdims = tuple(int(rng.integers(1, 9)) for _ in range(3))
        info = dict(B=B, C=C, cdims=cdims, ishape=ishape, oshape=oshape)
        src = synth.volume("v1", B, *ishape, C, rng)
        cage = np.stack([synth.alignet_cage(1, max(max(cdims), 2), rng)[0][:, :cdims[0], :cdims[1], :cdims[2]] for _ in range(B)]).astype(np.float32)
        if rng.random() < 0.3: cage = rng.uniform(-1.6, 1.6, cage.shape).astype(np.float32)
        go = synth.grad_output(B, *oshape, C, rng)
        ref_out, ref_gsrc, ref_gcage = cage_chain_oracle(src, cage, oshape, go)
        m = vnn.CageWarpVolumetric(*oshape)
        out = host(m.updateOutput([dev(src), dev(cage)]))
        gsrc, gcage = m.updateGradInput([dev(src), dev(cage)], dev(go))
        if not bits_equal(out, ref_out): fail(kind, "out", info)
        if max_rel(host(gsrc), ref_gsrc) > 1e-4: fail(kind, "gradSource", info)
        if max_rel(host(gcage), ref_gcage) > 1e-4: fail(kind, "gradCage", (info, max_rel(host(gcage), ref_gcage)))
        if C == 1:
            tgt = (rng.uniform(size=(B,) + oshape + (1,)) > 0.5).astype(np.float32)
            st = vnn.CageWarpSmoothL1Volumetric(keepOutput=True)
            loss, (_, gc) = st.forwardBackward([dev(src), dev(cage)], dev(tgt))
            ref_loss = oracle.smooth_l1_forward(ref_out, tgt, True)
            _, _, ref_gc = cage_chain_oracle(src, cage, oshape, oracle.smooth_l1_backward(ref_out, tgt, True))
            if not bits_equal(host(st.output), ref_out): fail(kind, "step out", info)
            if abs(float(loss) - ref_loss) > 1e-9 * max(1.0, abs(ref_loss)): fail(kind, "step loss", info)
            if not np.array_equal(st.iouCounts.cpu().numpy().astype(np.uint64), oracle.iou_counts(ref_out, tgt)): fail(kind, "step iou", info)
            if max_rel(host(gc), ref_gc) > 2e-4: fail(kind, "step gradCage", (info, max_rel(host(gc), ref_gc)))
    elif kind == "affine":
        oshape = tuple(max(2, s) for s in oshape)
        C = int(rng.integers(1, 5))
        info = dict(B=B, C=C, ishape=ishape, oshape=oshape)
        vol = synth.volume("v1", B, *ishape, C, rng)
        T = random_affine(B, rng)
        go = synth.grad_output(B, *oshape, C, rng)
        grid = host(vnn.VolumetricGridGenerator(*oshape).updateOutput(dev(T)))
        ref_out = oracle.sampler_forward(vol, grid)
        ref_gv, ref_gg = oracle.sampler_backward(vol, grid, go)
        m = vnn.AffineSamplerVolumetric(*oshape)
        out = host(m.updateOutput([dev(vol), dev(T)]))
        gv, gT = m.updateGradInput([dev(vol), dev(T)], dev(go))
        if not bits_equal(out, ref_out): fail(kind, "out", info)
        if max_rel(host(gv), ref_gv) > 1e-4: fail(kind, "gradVol", info)
        if max_rel(host(gT), oracle.gridgen_backward(ref_gg)) > 2e-5: fail(kind, "gradT", (info, max_rel(host(gT), oracle.gridgen_backward(ref_gg))))
    else:
        N = int(rng.integers(1, 4))
        dims = [int(rng.integers(2, 14)) for _ in range(N)]
        x = rng.standard_normal([B, N] + dims).astype(np.float32)
        cs = vnn.Cumsum(allowSingletonBatch=True)
        info = dict(B=B, dims=dims)
        if not bits_equal(host(cs.updateOutput(dev(x))), oracle.cumsum_forward(x)): fail(kind, "fwd", info)
        if not bits_equal(host(cs.updateGradInput(dev(x), dev(x))), oracle.cumsum_backward(x)): fail(kind, "bwd", info)
print(f"{n} cases {counts}, {bad} failures")
