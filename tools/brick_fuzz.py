#!/usr/bin/env python
"""tools/brick_fuzz.py [n] -- random shapes / cages through the brick kernels (VOLSTN_FAST_ARITH) against the oracle chain.
Developer stress run (the committed tests cover fixed shapes): prints the worst deviations and every case beyond tolerance."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from alignet_b200 import nn as vnn, synth, _abi
from oracle import oracle
from tests.test_gpu_rowwarp import cage_chain_oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1234)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
worst_out = worst_loss = 0.0
bad = 0
for it in range(n):
    B = int(rng.integers(1, 5))
    oshape = (int(rng.integers(2, 24)), int(rng.integers(2, 24)), int(rng.integers(2, 80)))
    ishape = oshape if rng.random() < 0.5 else (int(rng.integers(1, 24)), int(rng.integers(1, 24)), int(rng.integers(1, 70)))
    cdims = tuple(int(rng.integers(1, 10)) for _ in range(3))
    kind = rng.integers(0, 4)
    cage = np.stack([synth.alignet_cage(1, max(max(cdims), 2), rng, jitter=float(rng.choice([0.0, 0.05, 0.25])))[0][:, :cdims[0], :cdims[1], :cdims[2]]
                     for _ in range(B)]).astype(np.float32)
    if kind == 1: cage *= float(rng.uniform(0.5, 2.5))
    if kind == 2: cage = rng.uniform(-1.6, 1.6, cage.shape).astype(np.float32)
    if kind == 3: cage = np.clip(cage * 1.2, -1, 1).astype(np.float32)
    src = np.ascontiguousarray(synth.volume("v2", B, *[max(s, 2) for s in ishape], 1, rng)[:, :ishape[0], :ishape[1], :ishape[2]])
    tgt = (rng.uniform(size=(B,) + oshape + (1,)) > 0.6).astype(np.float32)
    go = rng.standard_normal((B,) + oshape + (1,)).astype(np.float32)
    ref_out, ref_gsrc, ref_gcage = cage_chain_oracle(src, cage, oshape, go)
    scale = max(1e-30, float(np.max(np.abs(ref_out)))) * max(1.0, max(ishape) / 64.0)
    m = vnn.CageWarpVolumetric(*oshape, fastArith=True)
    out = m.updateOutput([dev(src), dev(cage)]).cpu().numpy()
    gsrc, gcage = m.updateGradInput([dev(src), dev(cage)], dev(go))
    e_out = float(np.max(np.abs(out - ref_out))) / scale
    e_gs = float(np.max(np.abs(gsrc.cpu().numpy() - ref_gsrc))) / max(1e-30, float(np.max(np.abs(ref_gsrc))))
    st = vnn.CageWarpSmoothL1Volumetric(fastArith=True, keepOutput=True)
    loss, (_, gc) = st.forwardBackward([dev(src), dev(cage)], dev(tgt))
    ref_loss = oracle.smooth_l1_forward(ref_out, tgt, True)
    e_loss = abs(float(loss) - ref_loss) / max(1.0, abs(ref_loss))
    ok = np.isfinite(out).all() and np.isfinite(gcage.cpu().numpy()).all() and np.isfinite(gc.cpu().numpy()).all() and e_out <= 4e-5 and e_gs <= 2e-4 and e_loss <= 3e-6
    worst_out, worst_loss = max(worst_out, e_out), max(worst_loss, e_loss)
    if not ok:
        bad += 1
        print("BAD", it, dict(B=B, oshape=oshape, ishape=ishape, cdims=cdims, kind=int(kind)), e_out, e_gs, e_loss, flush=True)
print(f"{n} cases, {bad} beyond tolerance; worst out {worst_out:.2e} (of max|out| x extent/64), worst loss {worst_loss:.2e}")
