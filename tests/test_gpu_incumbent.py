"""The GPU incumbent: the reference's own CUDA kernels (stn3dtorch/BilinearSamplerVolumetric.cu) compiled
UNMODIFIED for sm_100a behind oracle/fake_thc (oracle/ref_cuda_shim.cu -> oracle/_ref/libcuvolstn_ref.so).

1. validate it against the reference's CPU code first (its shared-memory reductions carry no __syncwarp,
   SURVEY.md section 5) -- tolerance 1e-5 / 1e-4 of max|ref|, it sums in a different order;
2. time it beside this repo's kernels on the bench workload and record both (gpurun_out/r01_incumbent.json
   when that directory is writable).  The recompiled kernels are the baseline this repo has to beat.
"""
import json
import os

import numpy as np
import pytest
import torch

from alignet_b200 import nn as vnn, synth
from oracle import oracle
from tests._util import max_rel

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not oracle.have_ref_cuda(), reason="oracle/_ref/libcuvolstn_ref.so not built")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@needs_ref
@pytest.mark.parametrize("gk", ["g2", "g4"])
def test_incumbent_matches_cpu_reference(gk):
    rng = np.random.default_rng(31)
    B, N, C = 2, 16, 1
    vol = synth.volume("v1", B, N, N, N, C, rng)
    grid = synth.grid(gk, B, N, N, N, rng, csz=4)
    go = synth.grad_output(B, N, N, N, C, rng)
    tv, tg, tgo = (torch.from_numpy(x).cuda() for x in (vol, grid, go))
    out = torch.empty_like(tgo)
    oracle.ref_cuda_call(0, [tv, tg, out])
    assert max_rel(out.cpu().numpy(), oracle.sampler_forward(vol, grid)) <= 1e-5
    gv, gg = torch.zeros_like(tv), torch.zeros_like(tg)
    oracle.ref_cuda_call(1, [tv, tg, gv, gg, tgo])
    rgv, rgg = oracle.sampler_backward(vol, grid, go)
    assert max_rel(gg.cpu().numpy(), rgg) <= 1e-5
    assert max_rel(gv.cpu().numpy(), rgv) <= 1e-4


@needs_ref
def test_time_incumbent_beside_ours():
    N, B, C = 64, 64, 1
    rng = np.random.default_rng(7)
    nb = 8
    res = {"workload": f"{N}^3 x C={C}, batch {B}, fwd + gradVol zero-fill + bwd, device-resident, CUDA events, 5 reps"}
    for gk in ("g1", "g2"):
        vol = torch.from_numpy(synth.volume("v2", nb, N, N, N, C, rng)).cuda().repeat(B // nb, 1, 1, 1, 1).contiguous()
        grid = torch.from_numpy(synth.grid(gk, nb, N, N, N, rng)).cuda().repeat(B // nb, 1, 1, 1, 1).contiguous()
        go = torch.randn((B, N, N, N, C), device="cuda")
        out, gv, gg = torch.empty_like(go), torch.empty_like(vol), torch.empty_like(grid)
        m = vnn.BilinearSamplerVolumetric()

        def ours():
            m.updateOutput([vol, grid]); m.updateGradInput([vol, grid], go)

        def incumbent():
            oracle.ref_cuda_call(0, [vol, grid, out])
            gv.zero_(); gg.zero_()                      # the Lua module zeroes both (lua:101-104)
            oracle.ref_cuda_call(1, [vol, grid, gv, gg, go])

        times = {}
        for name, fn in (("ours", ours), ("incumbent", incumbent)):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record(); torch.cuda.synchronize()
            times[name] = e0.elapsed_time(e1) / 5
        V = B * N ** 3
        res[gk] = {"ours_ms": times["ours"], "incumbent_ms": times["incumbent"],
                   "ours_gvox_s": V / times["ours"] / 1e6, "incumbent_gvox_s": V / times["incumbent"] / 1e6,
                   "speedup": times["incumbent"] / times["ours"]}
        assert times["ours"] < times["incumbent"], res  # the recompiled sm_20 kernels are the bar to clear
    print(json.dumps(res))
    outdir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(outdir) and os.access(outdir, os.W_OK):
        json.dump(res, open(os.path.join(outdir, "r01_incumbent.json"), "w"), indent=1)
