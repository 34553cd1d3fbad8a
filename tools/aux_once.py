"""tools/aux_once.py -- run the grid generator (forward, backward) and Cumsum once at the bench size, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alignet_b200 import nn as vnn
N, B = 64, 64
gen = vnn.VolumetricGridGenerator(N, N, N)
T = torch.eye(4, device="cuda").repeat(B, 1, 1)
gg = torch.randn((B, N, N, N, 4), device="cuda")
for _ in range(3):
    gen.updateOutput(T); gen.updateGradInput(T, gg)
cs = vnn.Cumsum(); d = torch.rand((B, 3, 8, 8, 8), device="cuda")
for _ in range(3):
    cs.updateOutput(d); cs.updateGradInput(d, d)
torch.cuda.synchronize()
