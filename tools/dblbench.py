"""tools/dblbench.py -- float vs double (generic kernels) at 64^3 x 16: double costs 2.8x / 3.0x the float time for 2x the bytes."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from alignet_b200 import synth, nn as vnn
N, B = 64, 16
rng = np.random.default_rng(7)
for dt in (np.float32, np.float64):
    vol = torch.from_numpy(synth.volume("v2", B, N, N, N, 1, rng, dt)).cuda()
    grid = torch.from_numpy(synth.grid("g1", B, N, N, N, rng, dt)).cuda()
    go = torch.from_numpy(synth.grad_output(B, N, N, N, 1, rng, dt)).cuda()
    m = vnn.BilinearSamplerVolumetric()
    def timed(fn, k=5):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k * 1e3
    print(np.dtype(dt).name, "fwd %.1f us  bwd %.1f us" % (timed(lambda: m.updateOutput([vol, grid])), timed(lambda: m.updateGradInput([vol, grid], go))))
