"""Every DEVICE entry point is capturable in a CUDA graph: it launches on the caller's stream, allocates nothing, never
synchronises and never touches the legacy stream ("caller's stream, no implicit sync"; the B200 way to run a
launch-bound step is a graph -- bench.py replays the sampler step, the fused modules and the training step as graphs).
Each module call is captured once, replayed on fresh inputs of the same shape, and compared with the eager call."""
import numpy as np
import pytest
import torch

from alignet_b200 import nn as vnn, synth

pytestmark = pytest.mark.gpu


def _capture(fn):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()                                  # warm-up outside the capture (lazy attribute / module state)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    return g, out


def _same(a, b, tol=0.0):
    a, b = a.detach().float().cpu().numpy(), b.detach().float().cpu().numpy()
    if tol == 0.0:
        return np.array_equal(a, b)
    return float(np.max(np.abs(a - b))) <= tol * max(1e-30, float(np.max(np.abs(b))))


def test_device_entry_points_are_graph_capturable():
    rng = np.random.default_rng(3)
    B, N, csz = 3, 16, 4
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    vol, vol2 = dev(synth.volume("v2", B, N, N, N, 1, rng)), dev(synth.volume("v2", B, N, N, N, 1, rng))
    grid, grid2 = dev(synth.grid("g2", B, N, N, N, rng)), dev(synth.grid("g5", B, N, N, N, rng))
    go = dev(synth.grad_output(B, N, N, N, 1, rng))
    tgt = dev(synth.volume("v2", B, N, N, N, 1, rng))
    cage, cage2 = dev(synth.alignet_cage(B, csz, rng).astype(np.float32)), dev(synth.alignet_cage(B, csz, rng).astype(np.float32))
    T = torch.eye(4, device="cuda").repeat(B, 1, 1) + 0.05 * torch.randn(B, 4, 4, device="cuda")
    d = torch.rand(B, 3, csz, csz, csz, device="cuda")

    cases = []
    smp = vnn.BilinearSamplerVolumetric()
    cases.append(("sampler", lambda v, g: (smp.updateOutput([v, g]).clone(),) + tuple(x.clone() for x in smp.updateGradInput([v, g], go)),
                  (vol, grid), (vol2, grid2), 1e-4))
    gen = vnn.VolumetricGridGenerator(N, N, N)
    cases.append(("gridgen", lambda t, gg: (gen.updateOutput(t).clone(), gen.updateGradInput(t, gg).clone()), (T, grid), (T * 1.1, grid2), 1e-5))
    cs = vnn.Cumsum()
    cases.append(("cumsum", lambda x, y: (cs.updateOutput(x).clone(), cs.updateGradInput(x, y).clone()), (d, d), (d * 2, d + 1), 0.0))
    aff = vnn.AffineSamplerVolumetric(N, N, N)
    cases.append(("affine", lambda v, t: (aff.updateOutput([v, t]).clone(),) + tuple(x.clone() for x in aff.updateGradInput([v, t], go)),
                  (vol, T), (vol2, T * 0.9), 1e-4))
    for fast in (False, True):
        cw = vnn.CageWarpVolumetric(fastArith=fast)
        cases.append((f"cage fast={fast}", lambda v, c, cw=cw: (cw.updateOutput([v, c]).clone(),) + tuple(x.clone() for x in cw.updateGradInput([v, c], go)),
                      (vol, cage), (vol2, cage2), 1e-4))
        st = vnn.CageWarpSmoothL1Volumetric(fastArith=fast, keepOutput=True)

        def step(v, c, st=st):
            loss, (_, gc) = st.forwardBackward([v, c], tgt)
            return loss.clone().reshape(1), gc.clone(), st.output.clone(), st.iouCounts.clone()
        cases.append((f"cage step fast={fast}", step, (vol, cage), (vol2, cage2), 1e-4))
    ss = vnn.BilinearSamplerSmoothL1Volumetric(keepOutput=True)

    def sstep(v, g):
        loss, (_, gg) = ss.forwardBackward([v, g], tgt)
        return loss.clone().reshape(1), gg.clone(), ss.output.clone()
    cases.append(("sampler step", sstep, (vol, grid), (vol2, grid2), 0.0))

    for name, fn, args, args2, tol in cases:
        static = [a.clone() for a in args]
        g, outs = _capture(lambda: fn(*static))
        for new in (args2, args):                                   # replay on other inputs of the same shape, then on the first again
            for s, n in zip(static, new):
                s.copy_(n)
            g.replay()
            torch.cuda.synchronize()
            eager = fn(*[n.clone() for n in new])
            for o, e in zip(outs, eager):
                assert _same(o, e, tol), name
