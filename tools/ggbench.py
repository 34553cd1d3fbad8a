"""tools/ggbench.py -- nn.VolumetricGridGenerator backward at 64^3 x 64 (graph replay, L2 flushed); used to sweep the
tunables of gridgen.cu (profiles/r01_gridgen_backward.txt)."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch, statistics
from alignet_b200 import nn as vnn
N, B = 64, 64
gen = vnn.VolumetricGridGenerator(N, N, N)
T = torch.eye(4, device="cuda").repeat(B, 1, 1)
gg = torch.randn((B, N, N, N, 4), device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3): gen.updateGradInput(T, gg)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g): gen.updateGradInput(T, gg)
ts = []
for _ in range(15):
    flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
print("gridgen bwd %.1f us (median of 15, L2 flushed)" % statistics.median(ts))
