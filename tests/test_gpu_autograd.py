"""torch.autograd glue (alignet_b200/autograd.py): gradients that flow through the sampler and Cumsum inside a
PyTorch graph are the oracle's updateGradInput results."""
import numpy as np
import pytest
import torch

from alignet_b200 import autograd as vag, synth
from oracle import oracle
from tests._util import bits_equal, max_rel

pytestmark = pytest.mark.gpu


def test_sampler_function_matches_oracle_gradients():
    rng = np.random.default_rng(21)
    B, N, C = 2, 10, 2
    vol = synth.volume("v1", B, N, N, N, C, rng)
    grid = synth.grid("g2", B, N, N, N, rng, csz=4)
    go = synth.grad_output(B, N, N, N, C, rng)
    tv = torch.from_numpy(vol).cuda().requires_grad_(True)
    tg = torch.from_numpy(grid).cuda().requires_grad_(True)
    out = vag.bilinear_sampler_volumetric(tv, tg)
    assert bits_equal(out.detach().cpu().numpy(), oracle.sampler_forward(vol, grid))
    out.backward(torch.from_numpy(go).cuda())
    rgv, rgg = oracle.sampler_backward(vol, grid, go)
    assert bits_equal(tg.grad.cpu().numpy(), rgg)
    assert max_rel(tv.grad.cpu().numpy(), rgv) <= 1e-4


def test_cumsum_function_matches_oracle_gradients():
    rng = np.random.default_rng(22)
    x = rng.standard_normal((3, 3, 5, 6, 7)).astype(np.float32)
    g = rng.standard_normal(x.shape).astype(np.float32)
    tx = torch.from_numpy(x).cuda().requires_grad_(True)
    y = vag.cumsum(tx)
    assert bits_equal(y.detach().cpu().numpy(), oracle.cumsum_forward(x))
    y.backward(torch.from_numpy(g).cuda())
    assert bits_equal(tx.grad.cpu().numpy(), oracle.cumsum_backward(g))


def test_selective_backward_skips_unrequested_products():
    """vol without requires_grad -> onlyGrid (no gradVol computed), grid without -> onlyVolume; values unchanged."""
    rng = np.random.default_rng(23)
    B, N = 2, 9
    vol = synth.volume("v1", B, N, N, N, 1, rng)
    grid = synth.grid("g5", B, N, N, N, rng)
    go = synth.grad_output(B, N, N, N, 1, rng)
    rgv, rgg = oracle.sampler_backward(vol, grid, go)
    tv = torch.from_numpy(vol).cuda()
    tg = torch.from_numpy(grid).cuda().requires_grad_(True)
    vag.bilinear_sampler_volumetric(tv, tg).backward(torch.from_numpy(go).cuda())
    assert tv.grad is None and bits_equal(tg.grad.cpu().numpy(), rgg)
    tv2 = torch.from_numpy(vol).cuda().requires_grad_(True)
    tg2 = torch.from_numpy(grid).cuda()
    vag.bilinear_sampler_volumetric(tv2, tg2).backward(torch.from_numpy(go).cuda())
    assert tg2.grad is None and max_rel(tv2.grad.cpu().numpy(), rgv) <= 1e-4
