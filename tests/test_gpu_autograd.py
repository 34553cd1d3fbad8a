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


def test_fused_cage_warp_functions_match_the_module_chain_gradients():
    """cage_warp_volumetric and cage_warp_smooth_l1 inside a PyTorch graph against the un-fused Functions of this file"""
    from alignet_b200 import nn as vnn
    rng = np.random.default_rng(24)
    B, N, csz = 2, 18, 4
    src = torch.from_numpy(synth.volume("v2", B, N, N, N, 1, rng)).cuda()
    tgt = torch.from_numpy(synth.volume("v2", B, N, N, N, 1, rng)).cuda()
    cage0 = torch.from_numpy(synth.alignet_cage(B, csz, rng).astype(np.float32)).cuda()
    field_i = vnn.VolumetricGridGenerator(N, N, N).updateOutput(torch.eye(4, device="cuda").repeat(B, 1, 1))
    ones = torch.ones((B, csz, csz, csz, 1), device="cuda")
    # un-fused chain: cage up-sample -> source warp -> SmoothL1
    c1 = cage0.clone().requires_grad_(True)
    field = vag.bilinear_sampler_volumetric(torch.cat([c1.permute(0, 2, 3, 4, 1), ones], -1).contiguous(), field_i)
    est = vag.bilinear_sampler_volumetric(src, field)
    l1 = torch.nn.functional.smooth_l1_loss(est, tgt)
    l1.backward()
    # fused warp + torch criterion
    c2 = cage0.clone().requires_grad_(True)
    est2 = vag.cage_warp_volumetric(src, c2)
    assert torch.equal(est2.detach(), est.detach())
    l2 = torch.nn.functional.smooth_l1_loss(est2, tgt)
    l2.backward()
    # one-kernel step
    c3 = cage0.clone().requires_grad_(True)
    l3 = vag.cage_warp_smooth_l1(src, c3, tgt)
    (2.0 * l3).backward()
    scale = float(c1.grad.abs().max())
    assert float((c2.grad - c1.grad).abs().max()) <= 1e-4 * scale
    assert float((c3.grad - 2.0 * c1.grad).abs().max()) <= 2e-4 * scale
    assert abs(float(l3.detach()) - float(l1.detach())) <= 1e-6 * max(1.0, float(l1.detach()))
