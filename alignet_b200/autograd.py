"""torch.autograd glue over the nn mirror, so that the warp path can sit inside a PyTorch graph
(used by tools/train_step.py, the ALIGNet-3D training-step configuration of BASELINE.json).

Each Function calls the module's updateOutput in forward and updateGradInput in backward -- exactly
the calls torch7's nn.Sequential makes (train.lua:123-126) -- so everything below is the C ABI.
"""
from __future__ import annotations

import torch

from . import nn as vnn


class _Sampler(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vol, grid, module):
        ctx.module = module
        ctx.save_for_backward(vol, grid)
        return module.updateOutput([vol, grid]).clone()

    @staticmethod
    def backward(ctx, grad_out):
        vol, grid = ctx.saved_tensors
        # Selective backward: a product nobody asked for is not computed.  In ALIGNet
        # training the source volume is data (models/alignet_2d.lua:23-24 wraps it in nn.Identity), so the
        # source warp needs only gradGrid -- no zero-fill, no scatter -- and the cage up-sample's grid is the
        # constant identity field, so it needs only gradVol.
        need_vol, need_grid = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        m = ctx.module
        saved = (m.onlyGrid, m.onlyVolume)
        m.onlyGrid, m.onlyVolume = (not need_vol) and need_grid, need_vol and (not need_grid)
        try:
            gv, gg = m.updateGradInput([vol, grid], grad_out.contiguous())
        finally:
            m.onlyGrid, m.onlyVolume = saved
        return (gv.clone() if need_vol else None), (gg.clone() if need_grid else None), None


class _Cumsum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, module):
        ctx.module = module
        ctx.save_for_backward(x)
        return module.updateOutput(x)

    @staticmethod
    def backward(ctx, grad_out):
        (x,) = ctx.saved_tensors
        return ctx.module.updateGradInput(x, grad_out.contiguous()), None


def bilinear_sampler_volumetric(vol: torch.Tensor, grid: torch.Tensor, module: vnn.BilinearSamplerVolumetric = None):
    """out = sample(vol B x D x H x W x C, grid B x D x H x W x 4); differentiable in both arguments."""
    return _Sampler.apply(vol, grid, module or vnn.BilinearSamplerVolumetric())


def cumsum(x: torch.Tensor, module: vnn.Cumsum = None):
    """nn.Cumsum on B x N x d1..dN; differentiable."""
    return _Cumsum.apply(x, module or vnn.Cumsum())


class _CageWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vol, cage, module):
        ctx.module = module
        ctx.save_for_backward(vol, cage)
        return module.updateOutput([vol, cage]).clone()

    @staticmethod
    def backward(ctx, grad_out):
        vol, cage = ctx.saved_tensors
        need_vol, need_cage = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        m = ctx.module
        saved = m.onlyGrid
        m.onlyGrid = not need_vol   # the source is data in ALIGNet training (models/alignet_2d.lua:23-24)
        try:
            gv, gc = m.updateGradInput([vol, cage], grad_out.contiguous())
        finally:
            m.onlyGrid = saved
        return (gv.clone() if need_vol else None), (gc.clone() if need_cage else None), None


class _CageWarpSmoothL1(torch.autograd.Function):
    """loss = SmoothL1(CageWarp(vol, cage), target): forward AND backward happen in the forward call (one kernel);
    backward only scales the stored cage gradient by the incoming scalar."""

    @staticmethod
    def forward(ctx, vol, cage, target, module):
        loss, (gv, gc) = module.forwardBackward([vol, cage], target)
        ctx.save_for_backward(gc, gv)
        ctx.only_grid = module.onlyGrid
        return loss.to(vol.dtype)

    @staticmethod
    def backward(ctx, grad_loss):
        gc, gv = ctx.saved_tensors
        g = grad_loss.to(gc.dtype)
        need_vol = ctx.needs_input_grad[0] and not ctx.only_grid
        return (gv * g if need_vol else None), gc * g, None, None


def cage_warp_volumetric(vol: torch.Tensor, cage: torch.Tensor, module: vnn.CageWarpVolumetric = None):
    """out = sample(vol, up-sample(cage B x 3 x cD x cH x cW)); differentiable in both arguments (one kernel each way)."""
    return _CageWarp.apply(vol, cage, module or vnn.CageWarpVolumetric())


def cage_warp_smooth_l1(vol: torch.Tensor, cage: torch.Tensor, target: torch.Tensor,
                        module: vnn.CageWarpSmoothL1Volumetric = None):
    """SmoothL1Criterion(cage_warp(vol, cage), target) as a scalar, differentiable in the cage (and in the volume when the
    module was built with onlyGrid=False); the warp, the criterion and both backward passes are ONE kernel launch."""
    return _CageWarpSmoothL1.apply(vol, cage, target, module or vnn.CageWarpSmoothL1Volumetric())
