"""tools/upbench.py -- backward of the cage up-sample in the reference's own module graph (BilinearSamplerVolumetric with
a csz^3 x 4 volume under the 64^3 x 64 identity field): direct kernel vs the privatised kernel, cage sizes 4 / 8 / 12.
Results: profiles/r01_upsample_backward.txt."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from alignet_b200 import _abi, synth, nn as vnn
N, B = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 64)
rng = np.random.default_rng(7)
def timed(fn, k=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k * 1e3
field_i = vnn.VolumetricGridGenerator(N, N, N).updateOutput(torch.eye(4, device="cuda").repeat(B, 1, 1))
gfield = torch.randn((B, N, N, N, 4), device="cuda"); gfield[..., 3] = 0
for csz in (4, 8, 12, 16):
    cage = torch.from_numpy(synth.alignet_cage(8, csz, rng).astype(np.float32)).cuda().repeat(B // 8, 1, 1, 1, 1).contiguous()
    cage_vol = torch.ones((B, csz, csz, csz, 4), device="cuda"); cage_vol[..., :3] = cage.permute(0, 2, 3, 4, 1)
    ref = None
    for algo, nm in ((_abi.ALGO_DIRECT, "direct"), (_abi.ALGO_TILE, "tile"), (_abi.ALGO_AUTO, "auto")):
        for ov in (True, False):
            m = vnn.BilinearSamplerVolumetric(onlyVolume=ov, algo=algo)
            us = timed(lambda: m.updateGradInput([cage_vol, field_i], gfield))
            gv = m.updateGradInput([cage_vol, field_i], gfield)[0].clone()
            if ref is None: ref = gv
            err = float((gv - ref).abs().max() / ref.abs().max())
            print(f"csz={csz} up-sample bwd {'onlyVolume' if ov else 'full'} {nm}: {us:9.1f} us  rel-err vs direct {err:.2e}", flush=True)
