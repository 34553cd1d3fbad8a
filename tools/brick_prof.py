#!/usr/bin/env python
"""tools/brick_prof.py [--identity] [--exact] -- launches the brick kernels (cage warp) a few times for `ncu`:
    ncu --set full --import-source on -k regex:k_cage_brick -s 3 -c 3 -o gpurun_out/brick python tools/brick_prof.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from alignet_b200 import nn as vnn, synth

N, B, csz = 64, 64, 8
rng = np.random.default_rng(7)
nb = 8
rep = lambda x: torch.from_numpy(x).cuda().repeat(B // nb, *([1] * (x.ndim - 1))).contiguous()
vol = rep(synth.volume("v2", nb, N, N, N, 1, rng)); tgt = rep(synth.volume("v2", nb, N, N, N, 1, rng))
cage = rep(synth.alignet_cage(nb, csz, rng, jitter=0.0 if "--identity" in sys.argv else 0.25).astype(np.float32))
go = torch.randn((B, N, N, N, 1), device="cuda")
fa = "--exact" not in sys.argv
fast = vnn.CageWarpVolumetric(onlyGrid=True, fastArith=fa)
stp = vnn.CageWarpSmoothL1Volumetric(fastArith=fa)
for _ in range(2):
    fast.updateOutput([vol, cage]); fast.updateGradInput([vol, cage], go); stp.forwardBackward([vol, cage], tgt)
torch.cuda.synchronize()
