"""tools/stepbench.py -- the cage warp backward and the one-kernel training step, with and without the gradVol scatter
(used to choose the register caps of rowwarp.cu::row_min_ctas)."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from alignet_b200 import synth, nn as vnn
N, B, csz = 64, 64, 8
rng = np.random.default_rng(7)
rep = lambda x: torch.from_numpy(x).cuda().repeat(B // 8, *([1] * (x.ndim - 1))).contiguous()
vol, tgt = rep(synth.volume("v2", 8, N, N, N, 1, rng)), rep(synth.volume("v2", 8, N, N, N, 1, rng))
go = torch.randn((B, N, N, N, 1), device="cuda")
cage = rep(synth.alignet_cage(8, csz, rng).astype(np.float32))
def timed(fn, k=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k * 1e3
st, st2 = vnn.CageWarpSmoothL1Volumetric(), vnn.CageWarpSmoothL1Volumetric(onlyGrid=False)
cw, cw2 = vnn.CageWarpVolumetric(onlyGrid=True), vnn.CageWarpVolumetric()
print("step onlyGrid %.1f us | step full %.1f us | bwd onlyGrid %.1f us | bwd full %.1f us" % (
    timed(lambda: st.forwardBackward([vol, cage], tgt)), timed(lambda: st2.forwardBackward([vol, cage], tgt)),
    timed(lambda: cw.updateGradInput([vol, cage], go)), timed(lambda: cw2.updateGradInput([vol, cage], go))))
