"""one launch of the cage up-sample backward (small volume, automatic algorithm) per cage size, for ncu"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from alignet_b200 import synth, nn as vnn
N, B = 64, 64
rng = np.random.default_rng(7)
field_i = vnn.VolumetricGridGenerator(N, N, N).updateOutput(torch.eye(4, device="cuda").repeat(B, 1, 1))
gfield = torch.randn((B, N, N, N, 4), device="cuda"); gfield[..., 3] = 0
for csz in (4, 8):
    cage_vol = torch.ones((B, csz, csz, csz, 4), device="cuda")
    m = vnn.BilinearSamplerVolumetric(onlyVolume=True)
    m.updateGradInput([cage_vol, field_i], gfield)
torch.cuda.synchronize(); print("ok")
