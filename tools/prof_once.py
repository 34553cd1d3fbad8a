#!/usr/bin/env python
"""tools/prof_once.py -- launches each kernel of interest ONCE (for `ncu --set full`): packed vs row vs fused,
forward and backward, 64^3 x batch 64 (or --size/--batch), identity field / near-identity transforms / ALIGNet cage."""
import argparse, ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from alignet_b200 import _abi, synth
from alignet_b200 import nn as vnn

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=64); ap.add_argument("--batch", type=int, default=64); ap.add_argument("--csz", type=int, default=8)
ap.add_argument("--what", default="packed,row,affine,cage,step,cf")
ap.add_argument("--grid", default="g1", help="warp field family of the plain sampler launches (g1 identity, g2 ALIGNet-like, g3 uniform random)")
a = ap.parse_args()
N, B, csz = a.size, a.batch, a.csz
rng = np.random.default_rng(7)
nb = min(B, 8)
rep = lambda x: torch.from_numpy(x).cuda().repeat(B // nb, *([1] * (x.ndim - 1))).contiguous()
vol = rep(synth.volume("v2", nb, N, N, N, 1, rng)); go = torch.randn((B, N, N, N, 1), device="cuda")
grid = rep(synth.grid(a.grid, nb, N, N, N, rng))
T = torch.eye(4, device="cuda").repeat(B, 1, 1) + 0.02 * torch.randn(B, 4, 4, device="cuda")
cage = torch.from_numpy(synth.alignet_cage(nb, csz, rng).astype(np.float32)).cuda().repeat(B // nb, 1, 1, 1, 1).contiguous()
what = a.what.split(",")
for algo, nm in ((0, "packed"), (_abi.ALGO_ROW, "row")):
    if nm in what:
        for kw in ({}, {"onlyGrid": True}):
            m = vnn.BilinearSamplerVolumetric(algo=algo, **kw)
            if not kw: m.updateOutput([vol, grid])
            m.updateGradInput([vol, grid], go)
if "affine" in what:
    m = vnn.AffineSamplerVolumetric(N, N, N); m.updateOutput([vol, T]); m.updateGradInput([vol, T], go)
if "cage" in what:
    for og in (True, False):
        m = vnn.CageWarpVolumetric(onlyGrid=og)
        if og: m.updateOutput([vol, cage])
        m.updateGradInput([vol, cage], go)
if "step" in what:
    tgt = rep(synth.volume("v2", nb, N, N, N, 1, rng))
    vnn.CageWarpSmoothL1Volumetric().forwardBackward([vol, cage], tgt)
if "cf" in what:
    C, Bc = 4, max(1, B // 4)
    volc = torch.randn((Bc, C, N, N, N), device="cuda"); goc = torch.randn((Bc, C, N, N, N), device="cuda")
    gridc = rep(synth.grid("g2", nb, N, N, N, rng))[:Bc].permute(0, 4, 1, 2, 3).contiguous()
    m = vnn.BilinearSamplerVolumetric(layout="BCDHW"); m.updateOutput([volc, gridc]); m.updateGradInput([volc, gridc], goc)
torch.cuda.synchronize()
print("ok")
