"""Parity of the ROW kernels (alignet_b200/csrc/rowwarp.cu) -- SURVEY.md section 8, rows f1, f2, f3.

Every case is checked against the ORACLE CHAIN of the un-fused reference modules on the same inputs:

  f2  strided / planar tensors   oracle.sampler_* on contiguous copies of the permuted views
  f3  AffineSamplerVolumetric    grid generator -> oracle.sampler_*            (stn3dtorch/test.lua:15-25)
  f1  CageWarpVolumetric         identity base grid -> oracle.sampler_forward(cage volume) -> oracle.sampler_*(source)
                                 (models/alignet_2d.lua:136-147, alignet.lua:4-39 lifted to 3-D)

Tolerances (per tensor): out, gradGrid and the floor/mask decisions behind them bit-exact; gradVol and gradCage
(atomic accumulation order) max|d| <= 1e-4 * max|ref|; gradT (a reduction over all voxels) <= 1e-5 * max|ref|.
"""
import ctypes

import numpy as np
import pytest
import torch

from alignet_b200 import _abi, synth
from alignet_b200 import nn as vnn
from oracle import oracle
from tests._util import bits_equal, describe_mismatch, max_rel

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().contiguous().cpu().numpy()


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---------------------------------------------------------------------------------------------
# the row kernels on the packed layout (forced with VOLSTN_ALGO_ROW): same contract as the packed kernels
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("G", [4, 3])
@pytest.mark.parametrize("C", [1, 2, 3, 6])
@pytest.mark.parametrize("oshape", [(7, 10, 37), (5, 9, 20), (3, 5, 70)])
def test_row_kernels_on_packed_tensors(G, C, oshape):
    for gk in ("g1", "g2", "g3", "g4", "g5"):
        rng = np.random.default_rng(100 * G + 10 * C + ord(gk[1]) + oshape[2])
        B, ishape = 3, (9, 11, 13)
        vol = synth.volume("v1", B, *ishape, C, rng)
        grid = synth.grid(gk, B, *oshape, rng, G=G, csz=4)
        go = synth.grad_output(B, *oshape, C, rng)
        ref_out = oracle.sampler_forward(vol, grid, G)
        ref_gv, ref_gg = oracle.sampler_backward(vol, grid, go, G)
        m = vnn.BilinearSamplerVolumetric(gridWidth=G, algo=_abi.ALGO_ROW)
        out = host(m.updateOutput([dev(vol), dev(grid)]))
        assert bits_equal(out, ref_out), (gk, describe_mismatch(out, ref_out))
        gv, gg = m.updateGradInput([dev(vol), dev(grid)], dev(go))
        assert bits_equal(host(gg), ref_gg), (gk, describe_mismatch(host(gg), ref_gg))
        assert max_rel(host(gv), ref_gv) <= 1e-4, gk
        for kw in (dict(onlyGrid=True), dict(onlyVolume=True)):
            m2 = vnn.BilinearSamplerVolumetric(gridWidth=G, algo=_abi.ALGO_ROW, **kw)
            gv2, gg2 = m2.updateGradInput([dev(vol), dev(grid)], dev(go))
            if "onlyGrid" in kw:
                assert bits_equal(host(gg2), ref_gg) and float(gv2.abs().max()) == 0.0
            else:
                assert max_rel(host(gv2), ref_gv) <= 1e-4 and float(gg2.abs().max()) == 0.0


def test_row_kernels_special_values(golden):
    """non-finite and huge coordinates, exact +-1: the golden vectors of the reference's C code"""
    for (name, G, dt), c in golden.items():
        if dt != "float32":
            continue
        m = vnn.BilinearSamplerVolumetric(gridWidth=G, algo=_abi.ALGO_ROW)
        out = host(m.updateOutput([dev(c["vol"]), dev(c["grid"])]))
        assert bits_equal(out, c["out"]), (name, G, describe_mismatch(out, c["out"]))
        gv, gg = m.updateGradInput([dev(c["vol"]), dev(c["grid"])], dev(c["gradOut"]))
        assert bits_equal(host(gg), c["gradGrid"]), (name, G)
        assert np.array_equal(np.isnan(host(gv)), np.isnan(c["gradVol"])) and max_rel(host(gv), c["gradVol"]) <= 1e-4


# ---------------------------------------------------------------------------------------------
# f2: channel-first tensors in place (no nn.Transpose copies)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("G", [4, 3])
@pytest.mark.parametrize("C", [1, 2, 4])
def test_channel_first_layout(G, C):
    rng = np.random.default_rng(31 + G + C)
    B, ishape, oshape = 2, (6, 9, 12), (5, 8, 41)
    vol = synth.volume("v1", B, *ishape, C, rng)            # B D H W C
    for gk in ("g2", "g4"):
        grid = synth.grid(gk, B, *oshape, rng, G=G, csz=4)
        go = synth.grad_output(B, *oshape, C, rng)
        ref_out = oracle.sampler_forward(vol, grid, G)
        ref_gv, ref_gg = oracle.sampler_backward(vol, grid, go, G)
        cf = lambda a: dev(np.moveaxis(a, -1, 1))           # materialise B C D H W
        m = vnn.BilinearSamplerVolumetric(gridWidth=G, layout="BCDHW")
        out = m.updateOutput([cf(vol), cf(grid)])
        assert tuple(out.shape) == (B, C) + oshape and out.is_contiguous()
        assert bits_equal(host(out.permute(0, 2, 3, 4, 1)), ref_out), describe_mismatch(host(out.permute(0, 2, 3, 4, 1)), ref_out)
        gv, gg = m.updateGradInput([cf(vol), cf(grid)], cf(go))
        assert tuple(gv.shape) == (B, C) + ishape and tuple(gg.shape) == (B, G) + oshape
        assert bits_equal(host(gg.permute(0, 2, 3, 4, 1)), ref_gg)
        assert max_rel(host(gv.permute(0, 2, 3, 4, 1)), ref_gv) <= 1e-4


def test_projector_permutation_direct_abi():
    """stn3dtorch/test.lua:17,21: nn.Transpose({2,4},{4,5}) turns B x C x D x H x W into B x H x D x W x C, and
    nn.Transpose({4,5},{2,4}) undoes it on the output.  Here both are views handed to the C ABI."""
    rng = np.random.default_rng(41)
    B, C, D, H, W = 2, 2, 6, 7, 9
    vol_cf = rng.standard_normal((B, C, D, H, W)).astype(np.float32)
    view = lambda t: t.permute(0, 3, 2, 4, 1)               # B H D W C
    grid = synth.grid("g5", B, H, D, W, rng)
    go_cf = rng.standard_normal((B, C, D, H, W)).astype(np.float32)
    tv, tg, tgo = dev(vol_cf), dev(grid), dev(go_cf)
    out_cf = torch.empty((B, C, D, H, W), device="cuda")
    gv_cf = torch.empty_like(tv)
    gg = torch.empty_like(tg)
    lib, d = _abi.lib(), vnn._desc
    _abi.check(lib.volstn_b200_sampler_forward(0, 4, d(view(tv)), d(tg), d(view(out_cf)), 0, stream()), "fwd")
    _abi.check(lib.volstn_b200_sampler_backward(0, 4, d(view(tv)), d(tg), d(view(gv_cf)), d(gg), d(view(tgo)),
                                                _abi.BWD_ZERO_GRADVOL, stream()), "bwd")
    torch.cuda.synchronize()
    pv = lambda a: np.ascontiguousarray(np.transpose(a, (0, 3, 2, 4, 1)))
    ref_out = oracle.sampler_forward(pv(vol_cf), grid)
    ref_gv, ref_gg = oracle.sampler_backward(pv(vol_cf), grid, pv(go_cf))
    assert bits_equal(host(view(out_cf)), ref_out)
    assert bits_equal(host(gg), ref_gg)
    assert max_rel(host(view(gv_cf)), ref_gv) <= 1e-4


def test_channel_first_double_uses_generic_kernels():
    rng = np.random.default_rng(43)
    B, C = 2, 3
    vol = synth.volume("v1", B, 5, 6, 7, C, rng, np.float64)
    grid = synth.grid("g4", B, 4, 5, 9, rng, np.float64)
    go = synth.grad_output(B, 4, 5, 9, C, rng, np.float64)
    cf = lambda a: dev(np.moveaxis(a, -1, 1))
    m = vnn.BilinearSamplerVolumetric(layout="BCDHW")
    out = m.updateOutput([cf(vol), cf(grid)])
    assert bits_equal(host(out.permute(0, 2, 3, 4, 1)), oracle.sampler_forward(vol, grid))
    gv, gg = m.updateGradInput([cf(vol), cf(grid)], cf(go))
    ref_gv, ref_gg = oracle.sampler_backward(vol, grid, go)
    assert bits_equal(host(gg.permute(0, 2, 3, 4, 1)), ref_gg)
    assert max_rel(host(gv.permute(0, 2, 3, 4, 1)), ref_gv) <= 1e-12


# ---------------------------------------------------------------------------------------------
# f3: grid generator folded into the sampler
# ---------------------------------------------------------------------------------------------
def random_affine(B, rng, scale=0.3):
    T = np.tile(np.eye(4, dtype=np.float32), (B, 1, 1))
    q, _ = np.linalg.qr(rng.standard_normal((B, 3, 3)))
    T[:, :3, :3] = q.astype(np.float32)
    T[:, :3, 3] = (scale * rng.standard_normal((B, 3))).astype(np.float32)
    T[0] = np.eye(4, dtype=np.float32)                      # the identity: every coordinate exactly on the lattice
    return T


@pytest.mark.parametrize("C", [1, 3])
@pytest.mark.parametrize("oshape", [(6, 9, 40), (5, 7, 16), (64, 2, 3)])
def test_affine_sampler_matches_generator_then_sampler(C, oshape):
    rng = np.random.default_rng(51 + C + oshape[2])
    B, ishape = 3, (8, 9, 11)
    vol = synth.volume("v1", B, *ishape, C, rng)
    T = random_affine(B, rng)
    go = synth.grad_output(B, *oshape, C, rng)
    gen = vnn.VolumetricGridGenerator(*oshape)
    grid = host(gen.updateOutput(dev(T)))                   # the un-fused first stage (parity-tested on its own)
    assert np.abs(grid - oracle.gridgen_forward(T, *oshape)).max() <= 1e-6
    ref_out = oracle.sampler_forward(vol, grid)
    ref_gv, ref_gg = oracle.sampler_backward(vol, grid, go)
    ref_gT = oracle.gridgen_backward(ref_gg)
    m = vnn.AffineSamplerVolumetric(*oshape)
    out = host(m.updateOutput([dev(vol), dev(T)]))
    assert bits_equal(out, ref_out), describe_mismatch(out, ref_out)
    gv, gT = m.updateGradInput([dev(vol), dev(T)], dev(go))
    assert max_rel(host(gv), ref_gv) <= 1e-4
    assert max_rel(host(gT), ref_gT) <= 1e-5, (host(gT)[1], ref_gT[1])
    assert float(gT[:, 3].abs().max()) == 0.0               # gradGrid[..., 3] = 0 (...c:822)
    # channel-first volume in place, as in the projector
    m2 = vnn.AffineSamplerVolumetric(*oshape, layout="BCDHW")
    out2 = m2.updateOutput([dev(np.moveaxis(vol, -1, 1)), dev(T)])
    assert bits_equal(host(out2.permute(0, 2, 3, 4, 1)), ref_out)
    gv2, gT2 = m2.updateGradInput([dev(np.moveaxis(vol, -1, 1)), dev(T)], dev(np.moveaxis(go, -1, 1)))
    assert max_rel(host(gv2.permute(0, 2, 3, 4, 1)), ref_gv) <= 1e-4 and max_rel(host(gT2), ref_gT) <= 1e-5


def test_affine_sampler_projector_of_the_reference_test():
    """stn3dtorch/test.lua:87: the 90 degree rotation permutes and flips axes of the volume exactly"""
    N = 16
    rng = np.random.default_rng(7)
    vol = (rng.standard_normal((1, N, N, N, 1)) >= 0.5).astype(np.float32)      # test.lua:83
    T = np.zeros((1, 4, 4), np.float32)
    T[0, 0, 0] = 1
    T[0, 1, 2] = 1
    T[0, 2, 1] = -1
    T[0, 3, 3] = 1
    out = host(vnn.AffineSamplerVolumetric(N, N, N).updateOutput([dev(vol), dev(T)]))
    grid = host(vnn.VolumetricGridGenerator(N, N, N).updateOutput(dev(T)))
    assert bits_equal(out, oracle.sampler_forward(vol, grid))
    # out[z, y, x] = vol[z, x, N-1-y] up to the rounding of the lattice coordinates
    expect = np.flip(np.swapaxes(vol, 2, 3), axis=2)
    assert np.abs(out - expect).max() <= 1e-4


# ---------------------------------------------------------------------------------------------
# f1: cage up-sample + source warp
# ---------------------------------------------------------------------------------------------
def cage_chain_oracle(src, cage, oshape, go=None):
    """the un-fused module chain on the CPU oracle"""
    B = src.shape[0]
    base = oracle.gridgen_base(*oshape)                                        # fieldI (identity through the generator)
    field_i = np.ascontiguousarray(np.broadcast_to(base, (B,) + base.shape))
    cage_vol = np.ones(cage.shape[:1] + cage.shape[2:] + (4,), np.float32)     # channels-last, 4th component constant
    cage_vol[..., :3] = np.moveaxis(cage, 1, -1)
    field = oracle.sampler_forward(cage_vol, field_i)
    out = oracle.sampler_forward(src, field)
    if go is None:
        return out
    gsrc, gfield = oracle.sampler_backward(src, field, go)
    gcage_vol, _ = oracle.sampler_backward(cage_vol, field_i, gfield)
    return out, gsrc, np.ascontiguousarray(np.moveaxis(gcage_vol[..., :3], -1, 1))


@pytest.mark.parametrize("cdims", [(4, 4, 4), (8, 8, 8), (3, 5, 2), (12, 12, 12)])
@pytest.mark.parametrize("oshape", [(16, 16, 16), (7, 10, 37), (9, 6, 70)])
@pytest.mark.parametrize("C", [1, 2])
def test_cage_warp_matches_module_chain(cdims, oshape, C):
    rng = np.random.default_rng(61 + cdims[0] + oshape[2] + C)
    B = 3
    src = synth.volume("v1", B, *oshape, C, rng)
    cage = np.stack([synth.alignet_cage(1, max(cdims), rng)[0][:, :cdims[0], :cdims[1], :cdims[2]] for _ in range(B)]).astype(np.float32)
    cage[1] = rng.uniform(-1.6, 1.6, cage[1].shape).astype(np.float32)          # folded, partly outside the volume
    go = synth.grad_output(B, *oshape, C, rng)
    ref_out, ref_gsrc, ref_gcage = cage_chain_oracle(src, cage, oshape, go)
    m = vnn.CageWarpVolumetric()
    out = host(m.updateOutput([dev(src), dev(cage)]))
    assert bits_equal(out, ref_out), describe_mismatch(out, ref_out)
    gsrc, gcage = m.updateGradInput([dev(src), dev(cage)], dev(go))
    assert max_rel(host(gsrc), ref_gsrc) <= 1e-4
    assert max_rel(host(gcage), ref_gcage) <= 1e-4, describe_mismatch(host(gcage), ref_gcage)
    m2 = vnn.CageWarpVolumetric(onlyGrid=True)
    gsrc2, gcage2 = m2.updateGradInput([dev(src), dev(cage)], dev(go))
    assert float(gsrc2.abs().max()) == 0.0 and max_rel(host(gcage2), ref_gcage) <= 1e-4


def test_cage_warp_identity_cage_reproduces_the_source():
    """ALIGNet at initialisation (weight 0, bias del: models/alignet_2d.lua:77-80,118-119): the cage is
    linspace(-1, 1, csz) per axis and the warp is the identity."""
    N, csz, B = 32, 8, 2
    rng = np.random.default_rng(3)
    src = synth.volume("v2", B, N, N, N, 1, rng)
    ax = np.linspace(-1, 1, csz).astype(np.float32)
    cage = np.empty((B, 3, csz, csz, csz), np.float32)
    cage[:, 0] = ax[:, None, None]
    cage[:, 1] = ax[None, :, None]
    cage[:, 2] = ax[None, None, :]
    out = host(vnn.CageWarpVolumetric().updateOutput([dev(src), dev(cage)]))
    assert bits_equal(out, cage_chain_oracle(src, cage, (N, N, N)))
    assert np.abs(out - src).max() <= 2e-5


def test_cage_warp_output_resolution_differs_from_source():
    rng = np.random.default_rng(5)
    src = synth.volume("v1", 2, 6, 7, 8, 1, rng)
    cage = synth.alignet_cage(2, 4, rng).astype(np.float32)
    oshape = (5, 9, 33)
    go = synth.grad_output(2, *oshape, 1, rng)
    ref_out, ref_gsrc, ref_gcage = cage_chain_oracle(src, cage, oshape, go)
    m = vnn.CageWarpVolumetric(*oshape)
    assert bits_equal(host(m.updateOutput([dev(src), dev(cage)])), ref_out)
    gsrc, gcage = m.updateGradInput([dev(src), dev(cage)], dev(go))
    assert max_rel(host(gsrc), ref_gsrc) <= 1e-4 and max_rel(host(gcage), ref_gcage) <= 1e-4


def test_fused_modules_reject_what_they_do_not_support():
    lib, d = _abi.lib(), vnn._desc
    v = torch.zeros((1, 4, 4, 4, 1), device="cuda", dtype=torch.float64)
    o = torch.zeros((1, 4, 4, 4, 1), device="cuda", dtype=torch.float64)
    T = torch.eye(4, device="cuda", dtype=torch.float64).view(1, 4, 4)
    assert lib.volstn_b200_affine_sampler_forward(_abi.F64, T.data_ptr(), d(v), d(o), 0, stream()) == _abi.ERR_UNSUPPORTED
    vf, of = v.float(), o.float()
    dims = (ctypes.c_int64 * 3)(4, 4, 32)
    cage = torch.zeros((1, 3, 4, 4, 32), device="cuda")
    assert lib.volstn_b200_cage_warp_forward(_abi.F32, cage.data_ptr(), dims, d(vf), d(of), 0, stream()) == _abi.ERR_UNSUPPORTED
    o1 = torch.zeros((1, 1, 4, 4, 1), device="cuda")      # extent 1: the generator asserts depth > 1 (lua:6-8)
    assert lib.volstn_b200_affine_sampler_forward(_abi.F32, T.float().data_ptr(), d(vf), d(o1), 0, stream()) == _abi.ERR_SHAPE


# ---------------------------------------------------------------------------------------------
# the fused modules on HOST tensors (the loader-side augmentation of dataset_2d.lua:285-311 runs on CPU tensors)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("B", [1, 37])
def test_host_fused_modules_match_device_and_oracle(pinned, B):
    rng = np.random.default_rng(71 + B)
    oshape, C = (9, 12, 34), 1
    vol = synth.volume("v1", B, 8, 10, 12, C, rng)
    T = random_affine(B, rng)
    go = synth.grad_output(B, *oshape, C, rng)
    cage = synth.alignet_cage(B, 4, rng).astype(np.float32)
    t = lambda a: torch.from_numpy(a).pin_memory() if pinned else torch.from_numpy(a)
    # affine
    grid = oracle.gridgen_forward(T, *oshape, accumulate=np.float32)
    dev_grid = host(vnn.VolumetricGridGenerator(*oshape).updateOutput(dev(T)))
    ref_out = oracle.sampler_forward(vol, dev_grid)
    ref_gv, ref_gg = oracle.sampler_backward(vol, dev_grid, go)
    m = vnn.AffineSamplerVolumetric(*oshape)
    out = m.updateOutput([t(vol), t(T)])
    assert not out.is_cuda and bits_equal(out.numpy(), ref_out)
    gv, gT = m.updateGradInput([t(vol), t(T)], t(go))
    assert max_rel(gv.numpy(), ref_gv) <= 1e-4 and max_rel(gT.numpy(), oracle.gridgen_backward(ref_gg)) <= 1e-5
    assert np.abs(grid - dev_grid).max() <= 1e-6
    # cage
    src = synth.volume("v1", B, *oshape, C, rng)
    ref_out, ref_gsrc, ref_gcage = cage_chain_oracle(src, cage, oshape, go)
    mc = vnn.CageWarpVolumetric()
    out = mc.updateOutput([t(src), t(cage)])
    assert not out.is_cuda and bits_equal(out.numpy(), ref_out)
    gsrc, gcage = mc.updateGradInput([t(src), t(cage)], t(go))
    assert max_rel(gsrc.numpy(), ref_gsrc) <= 1e-4 and max_rel(gcage.numpy(), ref_gcage) <= 1e-4
    mo = vnn.CageWarpVolumetric(onlyGrid=True)
    gsrc2, gcage2 = mo.updateGradInput([t(src), t(cage)], t(go))
    assert float(gsrc2.abs().max()) == 0.0 and max_rel(gcage2.numpy(), ref_gcage) <= 1e-4
    # host tensors in fast arithmetic: the staged device copies are contiguous, so the brick kernels run underneath
    mf = vnn.CageWarpVolumetric(fastArith=True)
    outf = mf.updateOutput([t(src), t(cage)])
    assert not outf.is_cuda and max_rel(outf.numpy(), ref_out) <= 5e-5          # N(0, 1) volume: slopes of several units per voxel
    gsrcf, gcagef = mf.updateGradInput([t(src), t(cage)], t(go))
    assert max_rel(gsrcf.numpy(), ref_gsrc) <= 1e-4
    errf = np.abs(gcagef.numpy() - ref_gcage) / float(np.max(np.abs(ref_gcage)))
    assert float(np.mean(errf <= 1e-4)) >= 0.97 and float(errf.max()) <= 5e-2


# ---------------------------------------------------------------------------------------------
# f4: forward + SmoothL1Criterion + its gradient + IoU counts + backward in one pass
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("oshape,cdims", [((16, 16, 16), (4, 4, 4)), ((7, 10, 37), (8, 8, 8)), ((9, 6, 70), (3, 5, 2))])
@pytest.mark.parametrize("size_average", [True, False])
def test_cage_warp_smooth_l1_step(oshape, cdims, size_average):
    rng = np.random.default_rng(81 + oshape[2] + cdims[0])
    B = 3
    src = synth.volume("v2", B, *oshape, 1, rng)
    src[1] *= 3.5                                           # |out - target| > 1 somewhere: the linear branch and the clamp
    target = synth.volume("v2", B, *oshape, 1, rng)
    cage = np.stack([synth.alignet_cage(1, max(cdims), rng)[0][:, :cdims[0], :cdims[1], :cdims[2]] for _ in range(B)]).astype(np.float32)
    # the un-fused chain on the oracle
    ref_out = cage_chain_oracle(src, cage, oshape)
    ref_loss = oracle.smooth_l1_forward(ref_out, target, size_average)
    go = oracle.smooth_l1_backward(ref_out, target, size_average)
    _, ref_gsrc, ref_gcage = cage_chain_oracle(src, cage, oshape, go)
    ref_iou = oracle.iou_counts(ref_out, target)
    for only_grid in (True, False):
        m = vnn.CageWarpSmoothL1Volumetric(sizeAverage=size_average, onlyGrid=only_grid, keepOutput=True)
        loss, (gsrc, gcage) = m.forwardBackward([dev(src), dev(cage)], dev(target))
        assert bits_equal(host(m.output), ref_out)                                  # the forward value, bit for bit
        assert abs(float(loss) - ref_loss) <= 1e-9 * max(1.0, abs(ref_loss))        # double accumulation, order only
        assert np.array_equal(m.iouCounts.cpu().numpy().astype(np.uint64), ref_iou)  # exact integer counts
        assert max_rel(host(gcage), ref_gcage) <= 1e-4, describe_mismatch(host(gcage), ref_gcage)
        if only_grid:
            assert float(gsrc.abs().max()) == 0.0
        else:
            assert max_rel(host(gsrc), ref_gsrc) <= 1e-4
    # without keepOutput nothing but the gradients and the scalars leaves the kernel
    m = vnn.CageWarpSmoothL1Volumetric(sizeAverage=size_average)
    loss2, (_, gcage2) = m.forwardBackward([dev(src), dev(cage)], dev(target))
    assert abs(float(loss2) - ref_loss) <= 1e-9 * max(1.0, abs(ref_loss)) and max_rel(host(gcage2), ref_gcage) <= 1e-4
    ref_mean_iou = float(np.mean(ref_iou[:, 0].astype(np.float64) / ref_iou[:, 1].astype(np.float64)))
    assert abs(float(m.iou()) - ref_mean_iou) <= 1e-12


def test_smooth_l1_gradient_equals_chain_of_fused_modules():
    """the step's gradCage against CageWarpVolumetric fed with the criterion's gradient computed by torch"""
    rng = np.random.default_rng(91)
    B, N = 2, 24
    src, target = synth.volume("v2", B, N, N, N, 1, rng), synth.volume("v2", B, N, N, N, 1, rng)
    cage = synth.alignet_cage(B, 6, rng).astype(np.float32)
    ts, tc, tt = dev(src), dev(cage), dev(target)
    cw = vnn.CageWarpVolumetric(onlyGrid=True)
    out = cw.updateOutput([ts, tc]).clone().requires_grad_(True)
    l = torch.nn.functional.smooth_l1_loss(out, tt)
    l.backward()
    _, gc_chain = cw.updateGradInput([ts, tc], out.grad)
    step = vnn.CageWarpSmoothL1Volumetric()
    loss, (_, gc) = step.forwardBackward([ts, tc], tt)
    assert abs(float(loss) - float(l.detach())) <= 1e-6 * max(1.0, float(l.detach()))
    assert max_rel(host(gc), host(gc_chain)) <= 1e-4


def test_batches_beyond_one_launch():
    """more samples than gridDim.y allows (65535): the launchers walk the batch in chunks"""
    rng = np.random.default_rng(95)
    B = 66000
    src = synth.volume("v1", B, 2, 2, 3, 1, rng)
    go = synth.grad_output(B, 2, 2, 3, 1, rng)
    cage = rng.uniform(-1.2, 1.2, (B, 3, 2, 2, 2)).astype(np.float32)
    ref_out, ref_gsrc, ref_gcage = cage_chain_oracle(src, cage, (2, 2, 3), go)
    m = vnn.CageWarpVolumetric()
    assert bits_equal(host(m.updateOutput([dev(src), dev(cage)])), ref_out)
    gsrc, gcage = m.updateGradInput([dev(src), dev(cage)], dev(go))
    assert max_rel(host(gsrc), ref_gsrc) <= 1e-4 and max_rel(host(gcage), ref_gcage) <= 1e-4
    T = random_affine(B, rng)
    grid = host(vnn.VolumetricGridGenerator(2, 2, 3).updateOutput(dev(T)))
    ma = vnn.AffineSamplerVolumetric(2, 2, 3)
    assert bits_equal(host(ma.updateOutput([dev(src), dev(T)])), oracle.sampler_forward(src, grid))
    _, gT = ma.updateGradInput([dev(src), dev(T)], dev(go))
    assert max_rel(host(gT), oracle.gridgen_backward(oracle.sampler_backward(src, grid, go)[1])) <= 1e-5
    mr = vnn.BilinearSamplerVolumetric(algo=_abi.ALGO_ROW)
    assert bits_equal(host(mr.updateOutput([dev(src), dev(grid)])), oracle.sampler_forward(src, grid))
    step = vnn.CageWarpSmoothL1Volumetric(keepOutput=True)
    loss, _ = step.forwardBackward([dev(src), dev(cage)], dev(go))
    assert abs(float(loss) - oracle.smooth_l1_forward(ref_out, go)) <= 1e-9


# ---------------------------------------------------------------------------------------------
# BASELINE.json sizes: the fused modules against the chain of un-fused CUDA modules (each parity-tested against the
# oracle on its own) plus size-independent properties, and one sample against the oracle chain
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,B,csz", [(32, 32, 8), (64, 16, 8), (64, 8, 4)])
def test_fused_modules_at_full_size(N, B, csz):
    rng = np.random.default_rng(N + B + csz)
    src_np = synth.volume("v2", B, N, N, N, 1, rng)
    tgt_np = synth.volume("v2", B, N, N, N, 1, rng)
    cage_np = synth.alignet_cage(B, csz, rng).astype(np.float32)
    src, tgt, cage = dev(src_np), dev(tgt_np), dev(cage_np)
    go = dev(synth.grad_output(B, N, N, N, 1, rng))
    # the un-fused chain on the GPU
    field_i = vnn.VolumetricGridGenerator(N, N, N).updateOutput(torch.eye(4, device="cuda").repeat(B, 1, 1))
    cage_vol = torch.ones((B, csz, csz, csz, 4), device="cuda")
    cage_vol[..., :3] = cage.permute(0, 2, 3, 4, 1)
    up, warp = vnn.BilinearSamplerVolumetric(), vnn.BilinearSamplerVolumetric()
    field = up.updateOutput([cage_vol, field_i]).clone()
    ref_out = warp.updateOutput([src, field]).clone()
    ref_gsrc, gfield = [t.clone() for t in warp.updateGradInput([src, field], go)]
    ref_gcage = up.updateGradInput([cage_vol, field_i], gfield)[0][..., :3].permute(0, 4, 1, 2, 3).contiguous()
    # (1) the fused warp: same bits forward, same gradients within the accumulation tolerance
    m = vnn.CageWarpVolumetric()
    out = m.updateOutput([src, cage]).clone()
    assert torch.equal(out, ref_out)
    gsrc, gcage = m.updateGradInput([src, cage], go)
    assert float((gcage - ref_gcage).abs().max()) <= 1e-4 * float(ref_gcage.abs().max())
    assert float((gsrc - ref_gsrc).abs().max()) <= 1e-4 * float(ref_gsrc.abs().max())
    # (2) adjointness in the source: <warp(src), go> == <src, gradSource>
    lhs, rhs = float((out.double() * go.double()).sum()), float((gsrc.double() * src.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))
    # (3) the identity cage reproduces the source (ALIGNet at initialisation)
    ax = torch.linspace(-1, 1, csz, device="cuda")
    idc = torch.stack([ax.view(-1, 1, 1).expand(csz, csz, csz), ax.view(1, -1, 1).expand(csz, csz, csz),
                       ax.view(1, 1, -1).expand(csz, csz, csz)]).unsqueeze(0).repeat(B, 1, 1, 1, 1).contiguous()
    assert float((m.updateOutput([src, idc]) - src).abs().max()) <= 4e-7 * N
    # (4) the one-kernel step: forward bits, loss, exact IoU counts, cage gradient against torch's criterion on `ref_out`
    x = ref_out.clone().requires_grad_(True)
    ref_loss = torch.nn.functional.smooth_l1_loss(x.double(), tgt.double())
    ref_loss.backward()
    _, ref_gc2 = vnn.CageWarpVolumetric(onlyGrid=True).updateGradInput([src, cage], x.grad.float().contiguous())
    step = vnn.CageWarpSmoothL1Volumetric(keepOutput=True)
    loss, (_, gc2) = step.forwardBackward([src, cage], tgt)
    assert torch.equal(step.output, ref_out)
    assert abs(float(loss) - float(ref_loss.detach())) <= 1e-9 * max(1.0, float(ref_loss.detach()))
    inter = ((ref_out > 0.5) & (tgt > 0.5)).reshape(B, -1).sum(1)
    union = ((ref_out > 0.5) | (tgt > 0.5)).reshape(B, -1).sum(1)
    assert torch.equal(step.iouCounts[:, 0], inter) and torch.equal(step.iouCounts[:, 1], union)
    assert float((gc2 - ref_gc2).abs().max()) <= 2e-4 * float(ref_gc2.abs().max())
    # (5) first and last sample against the oracle chain, bit-exact
    for b in (0, B - 1):
        assert bits_equal(host(out[b:b + 1]), cage_chain_oracle(src_np[b:b + 1], cage_np[b:b + 1], (N, N, N)))
    # (6) grid generator folded into the sampler, same size: bits of the two-module chain
    T = dev(random_affine(B, rng))
    grid = vnn.VolumetricGridGenerator(N, N, N).updateOutput(T)
    assert torch.equal(vnn.AffineSamplerVolumetric(N, N, N).updateOutput([src, T]), warp.updateOutput([src, grid]))


@pytest.mark.parametrize("G", [4, 3])
@pytest.mark.parametrize("oshape", [(7, 10, 37), (6, 9, 20)])
def test_sampler_smooth_l1_step(G, oshape):
    """the one-kernel step with a grid tensor: forward bits, loss, exact IoU counts, gradGrid bit-exact against the
    oracle fed with the criterion's gradient, gradVol within 1e-4"""
    rng = np.random.default_rng(97 + G + oshape[2])
    B, ishape = 3, (8, 9, 11)
    vol = synth.volume("v2", B, *ishape, 1, rng)
    vol[1] *= 3.0
    target = synth.volume("v2", B, *oshape, 1, rng)
    for gk in ("g1", "g2", "g4"):
        grid = synth.grid(gk, B, *oshape, rng, G=G, csz=4)
        ref_out = oracle.sampler_forward(vol, grid, G)
        go = oracle.smooth_l1_backward(ref_out, target)
        ref_gv, ref_gg = oracle.sampler_backward(vol, grid, go, G)
        for only_grid in (True, False):
            m = vnn.BilinearSamplerSmoothL1Volumetric(gridWidth=G, onlyGrid=only_grid, keepOutput=True)
            loss, (gv, gg) = m.forwardBackward([dev(vol), dev(grid)], dev(target))
            assert bits_equal(host(m.output), ref_out), gk
            assert abs(float(loss) - oracle.smooth_l1_forward(ref_out, target)) <= 1e-9
            assert np.array_equal(m.iouCounts.cpu().numpy().astype(np.uint64), oracle.iou_counts(ref_out, target))
            assert bits_equal(host(gg), ref_gg), (gk, describe_mismatch(host(gg), ref_gg))
            if only_grid:
                assert float(gv.abs().max()) == 0.0
            else:
                assert max_rel(host(gv), ref_gv) <= 1e-4


def test_fused_modules_match_the_reference_chain_golden():
    """tests/golden/cage_warp_golden.npz: chains of the reference's own compiled C code (make_cage_golden.py)"""
    from tests.test_oracle_golden import load_cage_golden
    for name, c in load_cage_golden().items():
        src, cage, go = c["source"], c["cage"], c["gradOut"]
        oshape = c["out"].shape[1:4]
        m = vnn.CageWarpVolumetric(*oshape)
        out = host(m.updateOutput([dev(src), dev(cage)]))
        assert bits_equal(out, c["out"]), (name, describe_mismatch(out, c["out"]))
        gsrc, gcage = m.updateGradInput([dev(src), dev(cage)], dev(go))
        assert max_rel(host(gsrc), c["gradSource"]) <= 1e-4, name
        assert max_rel(host(gcage), c["gradCage"]) <= 1e-4, name
        # the same golden chains bound the tolerance mode (brick kernels: the golden tensors are contiguous, C = 1)
        if src.shape[-1] == 1:
            mf = vnn.CageWarpVolumetric(*oshape, fastArith=True)
            scale = max(1.0, max(src.shape[1:4]) / 64.0)
            assert max_rel(host(mf.updateOutput([dev(src), dev(cage)])), c["out"]) <= 5e-5 * scale, name
            gsf, gcf = mf.updateGradInput([dev(src), dev(cage)], dev(go))
            assert max_rel(host(gsf), c["gradSource"]) <= 1e-4, name
            errf = np.abs(host(gcf) - c["gradCage"]) / max(1e-30, float(np.max(np.abs(c["gradCage"]))))
            if name.startswith("identity"):
                # The identity cage puts EVERY sampling coordinate on an integer, where the warp has two one-sided
                # derivatives.  The reference's float arithmetic lands one ulp below the integer for about a third of
                # the positions (left derivative) and on it for the rest (right derivative); the separable arithmetic
                # lands exactly on it (always the right derivative).  On this binary volume the two differ by O(1), so
                # no agreement of gradCage is claimed here (measured: 24 % of the cells within 1e-4) -- only sanity.
                assert np.isfinite(host(gcf)).all()
            else:
                assert float(np.mean(errf <= 1e-4)) >= 0.9 and float(errf.max()) <= 8e-2, (name, float(np.mean(errf <= 1e-4)), float(errf.max()))
        if "step_loss" in c:
            step = vnn.CageWarpSmoothL1Volumetric(onlyGrid=False, keepOutput=True)  # output resolution = the target's
            loss, (gs, gc) = step.forwardBackward([dev(src), dev(cage)], dev(c["target"]))
            assert bits_equal(host(step.output), c["out"]), name
            assert abs(float(loss) - float(c["step_loss"])) <= 1e-9 * max(1.0, float(c["step_loss"]))
            assert np.array_equal(step.iouCounts.cpu().numpy().astype(np.uint64), c["step_iou"]), name
            assert max_rel(host(gc), c["step_gradCage"]) <= 1e-4 and max_rel(host(gs), c["step_gradSource"]) <= 1e-4, name


# ---------------------------------------------------------------------------------------------
# VOLSTN_FAST_ARITH: the cage kernels with the separable up-sample, FMA sums and the factored gradGrid.  A TOLERANCE
# mode (include/volstn_b200.h): same mathematics as the reference chain, different rounding.
# ---------------------------------------------------------------------------------------------
def _jittered_cages(B, cdims, rng):
    return np.stack([synth.alignet_cage(1, max(cdims), rng)[0][:, :cdims[0], :cdims[1], :cdims[2]] for _ in range(B)]).astype(np.float32)


@pytest.mark.parametrize("cdims", [(4, 4, 4), (8, 8, 8), (3, 5, 2), (12, 12, 12), (6, 7, 20)])
@pytest.mark.parametrize("oshape", [(16, 16, 16), (7, 10, 37), (9, 6, 70), (32, 32, 32)])
def test_fast_arith_field_is_the_reference_field_to_2e5(cdims, oshape):
    """The up-sampled field itself, read back through the source warp of a RAMP volume (a trilinear sampler reproduces a
    linear function): out = the field in voxel units.  Within 2e-5 of the reference chain's, relative to the extent
    (measured on a B200: <= 1.53e-5 = a few ulp of the normalised coordinate, the size of the reference field's own
    rounding error; header: VOLSTN_FAST_ARITH)."""
    rng = np.random.default_rng(7 + cdims[2] + oshape[2])
    B = 2
    cage = _jittered_cages(B, cdims, rng)
    cage[1] = np.clip(cage[1] * 1.1, -1.3, 1.3)                  # partly outside the volume: the zero-padding tiers
    worst = 0.0
    for axis in range(3):
        ramp = np.arange(oshape[axis], dtype=np.float32).reshape([-1 if i == axis else 1 for i in range(3)])
        src = np.ascontiguousarray(np.broadcast_to(ramp[None, ..., None], (B,) + oshape + (1,))).astype(np.float32)
        ref = cage_chain_oracle(src, cage, oshape)
        out = host(vnn.CageWarpVolumetric(fastArith=True).updateOutput([dev(src), dev(cage)]))
        inside = ref > 0                                         # away from the padding edge, where the value jumps to 0
        err = float(np.max(np.abs(out - ref)[inside])) / float(np.max(np.abs(ref)))
        worst = max(worst, err)
    assert worst <= 2e-5, worst


@pytest.mark.parametrize("cdims,oshape", [((4, 4, 4), (16, 16, 16)), ((8, 8, 8), (32, 32, 32)), ((3, 5, 2), (9, 6, 70)), ((8, 8, 8), (64, 64, 64))])
def test_fast_arith_cage_warp_within_tolerance_of_the_chain(cdims, oshape):
    rng = np.random.default_rng(17 + cdims[0] + oshape[2])
    B = 2
    src = synth.volume("v2", B, *oshape, 1, rng)                  # ALIGNet's binary occupancy volumes
    target = synth.volume("v2", B, *oshape, 1, rng)
    cage = _jittered_cages(B, cdims, rng)
    go = synth.grad_output(B, *oshape, 1, rng)
    ref_out, ref_gsrc, ref_gcage = cage_chain_oracle(src, cage, oshape, go)
    m = vnn.CageWarpVolumetric(fastArith=True)
    out = host(m.updateOutput([dev(src), dev(cage)]))
    assert max_rel(out, ref_out) <= 2e-5, max_rel(out, ref_out)      # measured <= 1.07e-5 (64^3)
    gsrc, gcage = m.updateGradInput([dev(src), dev(cage)], dev(go))
    assert max_rel(host(gsrc), ref_gsrc) <= 1e-4
    # gradCage: the warp is only C0 across the cell boundaries of the SOURCE volume, and on a binary volume the two
    # one-sided derivatives differ by O(1).  A voxel whose coordinate sits within rounding of a boundary may take the other
    # side (measured: a few voxels in 10^6, each moving one or two cage cells by ~1e-2 of max|gradCage|).  So: almost every
    # cell within the gradVol tolerance, and no cell off by more than a few per cent.
    def close_enough(g):
        err = np.abs(g - ref_gcage) / float(np.max(np.abs(ref_gcage)))
        return float(np.mean(err <= 1e-4)) >= 0.97 and float(err.max()) <= 5e-2, (float(np.mean(err <= 1e-4)), float(err.max()))
    ok, what = close_enough(host(gcage))
    assert ok, what
    gcage_og = host(vnn.CageWarpVolumetric(fastArith=True, onlyGrid=True).updateGradInput([dev(src), dev(cage)], dev(go))[1])
    ok, what = close_enough(gcage_og)
    assert ok, what
    # on a SMOOTH source volume the one-sided derivatives nearly coincide and the whole gradient agrees
    zz, yy, xx = np.meshgrid(*[np.linspace(-1, 1, n) for n in oshape], indexing="ij")
    smooth = np.exp(-2.0 * (zz ** 2 + yy ** 2 + xx ** 2)).astype(np.float32)
    src_s = np.ascontiguousarray(np.broadcast_to(smooth[None, ..., None], (B,) + oshape + (1,)))
    _, _, ref_gc_s = cage_chain_oracle(src_s, cage, oshape, go)
    gc_s = host(vnn.CageWarpVolumetric(fastArith=True, onlyGrid=True).updateGradInput([dev(src_s), dev(cage)], dev(go))[1])
    assert max_rel(gc_s, ref_gc_s) <= 2e-3, max_rel(gc_s, ref_gc_s)
    # the one-kernel training step
    ref_loss = oracle.smooth_l1_forward(ref_out, target, True)
    g_o = oracle.smooth_l1_backward(ref_out, target, True)
    _, _, ref_gc = cage_chain_oracle(src, cage, oshape, g_o)
    ref_iou = oracle.iou_counts(ref_out, target).astype(np.int64)
    st = vnn.CageWarpSmoothL1Volumetric(fastArith=True, keepOutput=True)
    loss, (_, gc) = st.forwardBackward([dev(src), dev(cage)], dev(target))
    assert max_rel(host(st.output), ref_out) <= 2e-5
    assert abs(float(loss) - ref_loss) <= 1e-6 * max(1.0, abs(ref_loss))
    err = np.abs(host(gc) - ref_gc) / float(np.max(np.abs(ref_gc)))
    assert float(np.mean(err <= 1e-4)) >= 0.97 and float(err.max()) <= 5e-2, (float(np.mean(err <= 1e-4)), float(err.max()))
    iou = st.iouCounts.cpu().numpy().astype(np.int64)
    assert np.all(np.abs(iou - ref_iou) <= np.maximum(2, ref_iou // 100000)), (iou, ref_iou)   # out == 0.5 to the last bit may flip
    # No claim is made about the source warp's floor indices in this mode: a field that moves by a few ulp changes the
    # output by slope x 1e-6 voxels whether or not a floor flips, so flips cannot be told from rounding in the values
    # -- which is the point: the value is continuous across a cell boundary.


@pytest.mark.parametrize("cdims,oshape,ishape", [((4, 4, 4), (16, 16, 16), None), ((1, 4, 3), (5, 9, 33), (6, 7, 8)),
                                                 ((3, 1, 2), (8, 5, 70), (4, 9, 5)), ((2, 2, 1), (6, 6, 6), (1, 5, 7)),
                                                 ((8, 8, 8), (64, 64, 64), (32, 32, 32)), ((5, 6, 7), (12, 40, 31), (128, 128, 9))])
def test_fast_arith_brick_kernels_padding_tiers_and_degenerate_cages(cdims, oshape, ishape):
    """The brick kernels (csrc/cagebrick.cu: contiguous tensors + VOLSTN_FAST_ARITH) where their two tiers and their
    table remapping are stressed: cages far outside the volume (general tier: clamped, shifted floor, predicated corners =
    the reference's zero padding ...c:542-564), folded cages, one-node cage axes, one-voxel source axes, a source
    resolution different from the output's (run-time and compile-time strides), ragged x slots.  Against the oracle
    chain at the tolerance of the mode, and against the row kernels in the same mode (VOLSTN_ALGO_ROW)."""
    rng = np.random.default_rng(23 + cdims[0] + oshape[2])
    B = 4
    ishape = ishape or oshape
    src = np.ascontiguousarray(synth.volume("v2", B, *[max(s, 2) for s in ishape], 1, rng)[:, :ishape[0], :ishape[1], :ishape[2]])
    tshape = (B,) + oshape + (1,)
    target = (rng.uniform(size=tshape) > 0.6).astype(np.float32)
    cage = np.stack([synth.alignet_cage(1, max(max(cdims), 2), rng)[0][:, :cdims[0], :cdims[1], :cdims[2]] for _ in range(B)]).astype(np.float32)
    cage[1] *= 2.5                                                              # most of the field far outside the volume
    cage[2] = rng.uniform(-1.6, 1.6, cage[2].shape).astype(np.float32)          # folded, partly outside
    cage[3] = np.clip(cage[3], -1.0, 1.0)                                       # faces hit exactly
    go = rng.standard_normal(tshape).astype(np.float32)
    ref_out, ref_gsrc, ref_gcage = cage_chain_oracle(src, cage, oshape, go)
    # a coordinate is known to a few ulp of the source extent: the tolerance of the mode scales with it beyond 64 voxels
    scale = max(1e-30, float(np.max(np.abs(ref_out)))) * max(1.0, max(ishape) / 64.0)
    for algo in (0, _abi.ALGO_ROW):
        m = vnn.CageWarpVolumetric(*oshape, fastArith=True, algo=algo)
        out = host(m.updateOutput([dev(src), dev(cage)]))
        assert np.isfinite(out).all()
        assert float(np.max(np.abs(out - ref_out))) <= 3e-5 * scale, (algo, float(np.max(np.abs(out - ref_out))) / scale)
        gsrc, gcage = m.updateGradInput([dev(src), dev(cage)], dev(go))
        assert max_rel(host(gsrc), ref_gsrc) <= 1e-4, algo
        err = np.abs(host(gcage) - ref_gcage) / max(1e-30, float(np.max(np.abs(ref_gcage))))
        assert float(np.mean(err <= 1e-4)) >= 0.9 and float(err.max()) <= 8e-2, (algo, float(np.mean(err <= 1e-4)), float(err.max()))
    # the two kernel families agree with each other much more closely than either has to agree with the chain
    a_out = host(vnn.CageWarpVolumetric(*oshape, fastArith=True).updateOutput([dev(src), dev(cage)]))
    r_out = host(vnn.CageWarpVolumetric(*oshape, fastArith=True, algo=_abi.ALGO_ROW).updateOutput([dev(src), dev(cage)]))
    assert float(np.max(np.abs(a_out - r_out))) <= 3e-5 * scale
    # the one-kernel step on the same inputs
    ref_loss = oracle.smooth_l1_forward(ref_out, target, True)
    g_o = oracle.smooth_l1_backward(ref_out, target, True)
    _, _, ref_gc = cage_chain_oracle(src, cage, oshape, g_o)
    ref_iou = oracle.iou_counts(ref_out, target).astype(np.int64)
    st = vnn.CageWarpSmoothL1Volumetric(fastArith=True, keepOutput=True)
    loss, (_, gc) = st.forwardBackward([dev(src), dev(cage)], dev(target))
    assert float(np.max(np.abs(host(st.output) - ref_out))) <= 3e-5 * scale
    assert abs(float(loss) - ref_loss) <= 2e-6 * max(1.0, abs(ref_loss))
    err = np.abs(host(gc) - ref_gc) / max(1e-30, float(np.max(np.abs(ref_gc))))
    assert float(np.mean(err <= 1e-4)) >= 0.9 and float(err.max()) <= 8e-2, (float(np.mean(err <= 1e-4)), float(err.max()))
    iou = st.iouCounts.cpu().numpy().astype(np.int64)
    assert np.all(np.abs(iou - ref_iou) <= np.maximum(2, ref_iou // 100000)), (iou, ref_iou)


def test_fast_arith_brick_kernels_survive_non_finite_cages():
    """NaN / Inf / huge cage nodes: the reference turns them into NaN outputs of the voxels they reach; the brick kernels
    clamp the coordinate (NaN -> far outside = zero padding).  Claimed: no fault, every voxel the bad nodes do NOT reach
    keeps its value, nothing non-finite leaks into the cage gradient of the other samples."""
    rng = np.random.default_rng(41)
    B, oshape, csz = 3, (16, 16, 40), 4
    src = synth.volume("v2", B, *oshape, 1, rng)
    cage = _jittered_cages(B, (csz, csz, csz), rng)
    clean = cage.copy()
    cage[1, 0, 1, 2, 1] = np.nan
    cage[1, 2, 3, 3, 3] = np.inf
    cage[1, 1, 0, 0, 0] = -3e38
    go = synth.grad_output(B, *oshape, 1, rng)
    m = vnn.CageWarpVolumetric(fastArith=True)
    out = host(m.updateOutput([dev(src), dev(cage)]))
    ref = host(vnn.CageWarpVolumetric(fastArith=True).updateOutput([dev(src), dev(clean)]))
    assert np.isfinite(out).all()
    assert np.array_equal(out[0], ref[0]) and np.array_equal(out[2], ref[2])
    _, gcage = m.updateGradInput([dev(src), dev(cage)], dev(go))
    _, gclean = vnn.CageWarpVolumetric(fastArith=True).updateGradInput([dev(src), dev(clean)], dev(go))
    g, gc = host(gcage), host(gclean)
    assert np.isfinite(g[0]).all() and np.isfinite(g[2]).all()
    assert max_rel(g[0], gc[0]) <= 1e-4 and max_rel(g[2], gc[2]) <= 1e-4


def test_fast_arith_identity_cage_reproduces_the_source():
    """ALIGNet at initialisation: every sampling coordinate sits ON a cell boundary, so the fast path may read the
    neighbouring cell with a weight within rounding of 0 or 1 -- the VALUE is the source voxel either way."""
    N, csz, B = 32, 8, 2
    rng = np.random.default_rng(3)
    src = synth.volume("v2", B, N, N, N, 1, rng)
    ax = np.linspace(-1, 1, csz).astype(np.float32)
    cage = np.empty((B, 3, csz, csz, csz), np.float32)
    cage[:, 0] = ax[:, None, None]
    cage[:, 1] = ax[None, :, None]
    cage[:, 2] = ax[None, None, :]
    out = host(vnn.CageWarpVolumetric(fastArith=True).updateOutput([dev(src), dev(cage)]))
    assert np.max(np.abs(out - src)) <= 2e-5
    ref = cage_chain_oracle(src, cage, (N, N, N))
    assert max_rel(out, ref) <= 2e-5


def test_fast_arith_is_opt_in_and_single_channel():
    """the default path stays bit-exact; multi-channel volumes ignore the flag (no fast kernels exist for them)"""
    rng = np.random.default_rng(5)
    B, oshape, cdims = 2, (9, 10, 37), (4, 4, 4)
    cage = _jittered_cages(B, cdims, rng)
    for C in (1, 2):
        src = synth.volume("v1", B, *oshape, C, rng)
        ref = cage_chain_oracle(src, cage, oshape)
        assert bits_equal(host(vnn.CageWarpVolumetric().updateOutput([dev(src), dev(cage)])), ref)
        fast = host(vnn.CageWarpVolumetric(fastArith=True).updateOutput([dev(src), dev(cage)]))
        if C == 2:
            assert bits_equal(fast, ref)
        else:
            assert max_rel(fast, ref) <= 5e-5          # N(0, 1) volume: slopes of several units per voxel
