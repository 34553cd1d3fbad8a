"""The drop-in luaT glue ON THE GPU: every lua_CFunction that lua/init.c (torch.FloatTensor / torch.DoubleTensor) and
lua/init.cu (torch.CudaTensor) register is called through the fake Lua stack of tests/luat_harness with tensors in the
reference's slots (self, 2..6: generic/BilinearSamplerVolumetric.c:444-446, 587-591) and compared with the REFERENCE'S
OWN stn3dtorch/init.c driven through the very same harness (oracle/_ref/libvolstn_refinit.so: init.c + the generic
file compiled unmodified).  Tolerances: out and gradGrid bit-exact, gradInputImages <= 1e-4 * max|ref| (atomics)."""
import os

import numpy as np
import pytest
import torch

from alignet_b200 import _abi, synth
from oracle import oracle
from tests._util import bits_equal, describe_mismatch, max_rel
from tests.luat_harness import luat

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def libs():
    import importlib.util
    spec = importlib.util.spec_from_file_location("build_glue", os.path.join(luat.ROOT, "lua", "build_glue.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    ours = luat.Harness(luat.DROPIN_CPU)
    cuda = luat.Harness(luat.DROPIN_CUDA)
    ref = luat.Harness(luat.REFERENCE_CPU) if os.path.exists(luat.REFERENCE_CPU) else None
    yield ours, cuda, ref
    for h in (ours, cuda, ref):
        if h is not None:
            h.close()


def _reference(ref, name, *args):
    """the reference's own lua_CFunction through the same stack; where oracle/_ref did not travel, the pinned oracle"""
    if ref is not None:
        assert ref.call(name, *args) == 1
        return
    G = 4 if "Volumetric" in name else 3
    if name.endswith("updateOutput"):
        vol, grid, out = args[:3]
        oracle.sampler_forward(vol, grid, G, out=out)
    else:
        vol, grid, gv, gg, go = args[:5]
        _, g = oracle.sampler_backward(vol, grid, go, G, grad_vol=gv)
        gg[...] = g


CASES = [("g4", 3, (6, 7, 8), (5, 6, 9), 2), ("g2", 4, (12, 12, 12), (12, 12, 12), 1), ("g3", 2, (5, 9, 33), (7, 8, 40), 4), ("g1", 2, (16, 16, 16), (16, 16, 16), 3)]


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("name,G", [("BilinearSamplerVolumetric", 4), ("BilinearSamplerBHWD", 3)])
def test_host_tensor_functions_equal_the_reference_init_c(libs, dt, name, G):
    ours, _, ref = libs
    for gk, B, ishape, oshape, C in CASES:
        rng = np.random.default_rng(hash((gk, B, C, G)) % 2 ** 31)
        vol = synth.volume("v1", B, *ishape, C, rng, dt)
        grid = synth.grid(gk, B, *oshape, rng, dt, G=G, csz=4)
        go = synth.grad_output(B, *oshape, C, rng, dt)
        # updateOutput(self, inputImages, grids, output)
        out_r, out_o = np.full((B, *oshape, C), 7, dt), np.full((B, *oshape, C), -7, dt)
        _reference(ref, name + "_updateOutput", vol, grid, out_r)
        assert ours.call(name + "_updateOutput", vol, grid, out_o) == 1
        assert bits_equal(out_o, out_r), (gk, describe_mismatch(out_o, out_r))
        # updateGradInput(self, inputImages, grids, gradInputImages, gradGrids, gradOutput): Lua zeroed both (lua:101-104)
        gv_r, gg_r, gv_o, gg_o = np.zeros_like(vol), np.zeros_like(grid), np.zeros_like(vol), np.zeros_like(grid)
        _reference(ref, name + "_updateGradInput", vol, grid, gv_r, gg_r, go)
        assert ours.call(name + "_updateGradInput", vol, grid, gv_o, gg_o, go) == 1
        assert bits_equal(gg_o, gg_r), (gk, describe_mismatch(gg_o, gg_r))
        assert max_rel(gv_o, gv_r) <= (1e-4 if dt == np.float32 else 1e-12), gk
        # ... and WITHOUT the Lua zero-fill both accumulate into gradInputImages (...c:737) and assign gradGrids (...c:819-822)
        init = rng.standard_normal(vol.shape).astype(dt)
        gv_r, gv_o = init.copy(), init.copy()
        gg_r, gg_o = np.full_like(grid, 3), np.full_like(grid, -3)
        _reference(ref, name + "_updateGradInput", vol, grid, gv_r, gg_r, go)
        assert ours.call(name + "_updateGradInput", vol, grid, gv_o, gg_o, go) == 1
        assert bits_equal(gg_o, gg_r) and max_rel(gv_o, gv_r) <= (1e-4 if dt == np.float32 else 1e-12), gk
        # the optional trailing number (the reference's unused focal_length slot, ...c:592) carries flags for the drop-in
        # Lua file: VOLSTN_BWD_ZERO_GRADVOL = the module's own :zero() moved into the call
        gv_z = init.copy()
        assert ours.call(name + "_updateGradInput", vol, grid, gv_z, gg_o, go, _abi.BWD_ZERO_GRADVOL) == 1
        gv_0 = np.zeros_like(vol)
        _reference(ref, name + "_updateGradInput", vol, grid, gv_0, gg_r, go)
        assert max_rel(gv_z, gv_0) <= (1e-4 if dt == np.float32 else 1e-12), gk


def test_host_strided_views_equal_the_reference_init_c(libs):
    """the reference honours the strides of dims 0-3 (...c:460-473, 607-630); so does the drop-in, without copies"""
    ours, _, ref = libs
    rng = np.random.default_rng(77)
    B, C = 3, 2
    vol = synth.volume("v1", B, 6, 7, 8, C, rng)
    big, gobig = synth.grid("g4", B, 6, 8, 20, rng), synth.grad_output(B, 6, 8, 20, C, rng)
    obig_r, obig_o = np.full_like(gobig, 9), np.full_like(gobig, 9)
    ix = (slice(None), slice(1, 5), slice(None), slice(2, 18, 3))
    grid, go = big[ix], gobig[ix]
    _reference(ref, "BilinearSamplerVolumetric_updateOutput", vol, grid, obig_r[ix])
    ours.call("BilinearSamplerVolumetric_updateOutput", vol, grid, obig_o[ix])
    assert bits_equal(obig_o, obig_r)                         # the view's elements equal, the rest of the buffer untouched
    gv_r, gv_o, ggb_r, ggb_o = np.zeros_like(vol), np.zeros_like(vol), np.full_like(big, 5), np.full_like(big, 5)
    _reference(ref, "BilinearSamplerVolumetric_updateGradInput", vol, grid, gv_r, ggb_r[ix], go)
    ours.call("BilinearSamplerVolumetric_updateGradInput", vol, grid, gv_o, ggb_o[ix], go)
    assert bits_equal(ggb_o, ggb_r) and max_rel(gv_o, gv_r) <= 1e-4


@pytest.mark.parametrize("name,G", [("BilinearSamplerVolumetric", 4), ("BilinearSamplerBHWD", 3)])
def test_cuda_tensor_functions_equal_the_reference_init_c(libs, name, G):
    """lua/init.cu: torch.CudaTensor arguments, launched on THC's current stream (...cu:706) -- compared with the
    reference's CPU init.c on the same data (the parity oracle is the CPU path, SURVEY 8 c)."""
    _, cuda, ref = libs
    st = torch.cuda.Stream()
    for gk, B, ishape, oshape, C in CASES:
        rng = np.random.default_rng(hash((gk, B, C, G, 1)) % 2 ** 31)
        vol = synth.volume("v1", B, *ishape, C, rng)
        grid = synth.grid(gk, B, *oshape, rng, G=G, csz=4)
        go = synth.grad_output(B, *oshape, C, rng)
        out_r, gv_r, gg_r = np.zeros((B, *oshape, C), np.float32), np.zeros_like(vol), np.zeros_like(grid)
        _reference(ref, name + "_updateOutput", vol, grid, out_r)
        _reference(ref, name + "_updateGradInput", vol, grid, gv_r, gg_r, go)
        with torch.cuda.stream(st):
            tv, tg, tgo = torch.from_numpy(vol).cuda(), torch.from_numpy(grid).cuda(), torch.from_numpy(go).cuda()
            out = torch.full((B, *oshape, C), 7.0, device="cuda")
            gv, gg = torch.zeros_like(tv), torch.full_like(tg, 7.0)
            n0 = _abi.launch_count()
            assert cuda.call(name + "_updateOutput", tv, tg, out, stream=st.cuda_stream) == 1
            assert cuda.call(name + "_updateGradInput", tv, tg, gv, gg, tgo, stream=st.cuda_stream) == 1
            assert _abi.launch_count() - n0 == 2      # one kernel each, nothing else
        st.synchronize()
        assert bits_equal(out.cpu().numpy(), out_r) and bits_equal(gg.cpu().numpy(), gg_r), gk
        assert max_rel(gv.cpu().numpy(), gv_r) <= 1e-4, gk
    with pytest.raises(luat.LuaError, match="torch.CudaTensor expected"):
        cuda.call("BilinearSamplerVolumetric_updateOutput", torch.zeros(1, 2, 2, 2, 1, device="cuda"), np.zeros((1, 2, 2, 2, 4), np.float32),
                  torch.zeros(1, 2, 2, 2, 1, device="cuda"))


def test_cumsum_and_grid_generator_functions(libs):
    """the four NEW native functions (Cumsum.lua / VolumetricGridGenerator.lua are pure Lua over TH in the reference):
    against the numpy restatements of oracle/oracle.py -- parity UNPINNED for these two modules."""
    ours, cuda, _ = libs
    rng = np.random.default_rng(5)
    for dt in (np.float32, np.float64):
        x = rng.standard_normal((6, 3, 8, 7, 9)).astype(dt)
        y, gx = np.empty_like(x), np.empty_like(x)
        assert ours.call("Cumsum_updateOutput", x, y) == 1 and ours.call("Cumsum_updateGradInput", x, gx) == 1
        if dt == np.float32:      # float: accumulated in double like TH's accreal, one rounding per element -> the same bits
            assert bits_equal(y, oracle.cumsum_forward(x)) and bits_equal(gx, oracle.cumsum_backward(x))
        else:                     # double: the warp scan adds in tree order
            assert max_rel(y, oracle.cumsum_forward(x)) <= 1e-14 and max_rel(gx, oracle.cumsum_backward(x)) <= 1e-14
        T = rng.standard_normal((4, 4, 4)).astype(dt)
        out = np.empty((4, 9, 10, 11, 4), dt)
        assert ours.call("VolumetricGridGenerator_updateOutput", T, out) == 1
        assert np.max(np.abs(out - oracle.gridgen_forward(T, 9, 10, 11))) <= 1e-6
        gg = rng.standard_normal(out.shape).astype(dt)
        gT = np.empty_like(T)
        assert ours.call("VolumetricGridGenerator_updateGradInput", gg, gT) == 1
        assert max_rel(gT, oracle.gridgen_backward(gg)) <= 1e-5
    # CudaTensor twins; the generator's backward takes a module-owned workspace tensor that the C function resizes
    x = rng.standard_normal((6, 3, 8, 8, 8)).astype(np.float32)
    tx, ty = torch.from_numpy(x).cuda(), torch.empty(x.shape, device="cuda")
    assert cuda.call("Cumsum_updateOutput", tx, ty) == 1
    assert bits_equal(ty.cpu().numpy(), oracle.cumsum_forward(x))
    assert cuda.call("Cumsum_updateGradInput", tx, ty) == 1
    assert bits_equal(ty.cpu().numpy(), oracle.cumsum_backward(x))
    T = rng.standard_normal((5, 4, 4)).astype(np.float32)
    out = torch.empty((5, 12, 13, 14, 4), device="cuda")
    assert cuda.call("VolumetricGridGenerator_updateOutput", torch.from_numpy(T).cuda(), out) == 1
    assert np.max(np.abs(out.cpu().numpy() - oracle.gridgen_forward(T, 12, 13, 14))) <= 1e-6
    gg = torch.randn(out.shape, device="cuda")
    gT, ws = torch.empty((5, 4, 4), device="cuda"), torch.empty(1 << 20, device="cuda")
    assert cuda.call("VolumetricGridGenerator_updateGradInput", gg, gT, luat.Tensor(ws)) == 1
    assert max_rel(gT.cpu().numpy(), oracle.gridgen_backward(gg.cpu().numpy())) <= 1e-5
    small = luat.Tensor(torch.empty(4, device="cuda"))
    with pytest.raises(luat.LuaError, match="workspace tensor too small"):
        cuda.call("VolumetricGridGenerator_updateGradInput", gg, gT, small)


def test_fused_cage_warp_functions(libs):
    """lua/CageWarpVolumetric.lua / lua/CageWarpSmoothL1Volumetric.lua call three more functions of libcuvolstn; each is
    driven through the fake stack with the module's slots and compared with the oracle chain of the un-fused reference
    modules (bit-exact `out`; gradients at the tolerances of tests/test_gpu_rowwarp.py) -- in both arithmetic modes."""
    from tests.test_gpu_rowwarp import cage_chain_oracle
    _, cuda, _ = libs
    rng = np.random.default_rng(19)
    B, oshape, csz = 3, (9, 10, 37), 4
    src = synth.volume("v2", B, *oshape, 1, rng)
    tgt = synth.volume("v2", B, *oshape, 1, rng)
    go = synth.grad_output(B, *oshape, 1, rng)
    cage = np.stack([synth.alignet_cage(1, csz, rng)[0] for _ in range(B)]).astype(np.float32)
    ref_out, ref_gsrc, ref_gcage = cage_chain_oracle(src, cage, oshape, go)
    ts, tc, tt, tgo = (torch.from_numpy(x).cuda() for x in (src, cage, tgt, go))
    ONLY_GRID, ZERO, FAST = _abi.BWD_ONLY_GRID, _abi.BWD_ZERO_GRADVOL, _abi.FAST_ARITH
    for fast in (0, FAST):
        out = torch.full((B, *oshape, 1), 7.0, device="cuda")
        assert cuda.call("CageWarpVolumetric_updateOutput", ts, tc, out, ZERO | fast) == 1
        if fast:
            assert max_rel(out.cpu().numpy(), ref_out) <= 2e-5
        else:
            assert bits_equal(out.cpu().numpy(), ref_out)
        gs, gc = torch.full_like(ts, 3.0), torch.full_like(tc, 3.0)
        assert cuda.call("CageWarpVolumetric_updateGradInput", ts, tc, gs, gc, tgo, ZERO | fast) == 1
        assert max_rel(gs.cpu().numpy(), ref_gsrc) <= 1e-4
        err = np.abs(gc.cpu().numpy() - ref_gcage) / float(np.max(np.abs(ref_gcage)))
        assert float(err.max()) <= (5e-2 if fast else 1e-4) and float(np.mean(err <= 1e-4)) >= 0.97
        # onlyGrid: gradSource is an EMPTY tensor in the Lua module and is not touched
        empty = torch.empty(0, device="cuda")
        gc2 = torch.full_like(tc, 5.0)
        assert cuda.call("CageWarpVolumetric_updateGradInput", ts, tc, empty, gc2, tgo, ZERO | ONLY_GRID | fast) == 1
        assert max_rel(gc2.cpu().numpy(), gc.cpu().numpy()) <= 1e-5
        # the one-kernel training step: workspace (device scratch) and stats (host DoubleTensor B x 3) are module-owned buffers
        n = int(np.prod(tgt.shape))
        ref_loss = oracle.smooth_l1_forward(ref_out, tgt, True)
        _, _, ref_gc = cage_chain_oracle(src, cage, oshape, oracle.smooth_l1_backward(ref_out, tgt, True))
        ref_iou = oracle.iou_counts(ref_out, tgt).astype(np.int64)
        stats = np.full((B, 3), -1.0, np.float64)
        ws = luat.Tensor(torch.empty(64, device="cuda"))
        kept, gcs = torch.full((B, *oshape, 1), 9.0, device="cuda"), torch.full_like(tc, 9.0)
        scale = float(np.float32(1.0 / np.float32(n)))
        assert cuda.call("CageWarpSmoothL1Volumetric_forwardBackward", ts, tc, tt, kept, empty, gcs, ws, stats, scale,
                         ZERO | ONLY_GRID | fast) == 1
        loss = stats[:, 0].sum() / n
        assert abs(loss - ref_loss) <= (1e-6 if fast else 1e-9) * max(1.0, abs(ref_loss))
        iou = stats[:, 1:].astype(np.int64)
        assert np.all(np.abs(iou - ref_iou) <= (2 if fast else 0)), (iou, ref_iou)
        if fast:
            assert max_rel(kept.cpu().numpy(), ref_out) <= 2e-5
        else:
            assert bits_equal(kept.cpu().numpy(), ref_out)
        err = np.abs(gcs.cpu().numpy() - ref_gc) / float(np.max(np.abs(ref_gc)))
        assert float(err.max()) <= (5e-2 if fast else 2e-4) and float(np.mean(err <= 2e-4)) >= 0.97
        # keepOutput = false: an empty output tensor
        stats2 = np.zeros((B, 3))
        assert cuda.call("CageWarpSmoothL1Volumetric_forwardBackward", ts, tc, tt, empty, empty, gcs, ws, stats2, scale,
                         ZERO | ONLY_GRID | fast) == 1
        assert np.array_equal(stats2[:, 1:], stats[:, 1:]) and np.allclose(stats2[:, 0], stats[:, 0], rtol=1e-12, atol=0)   # double atomics: order
    with pytest.raises(luat.LuaError, match="cage must be"):
        cuda.call("CageWarpVolumetric_updateOutput", ts, tc[:, :2].contiguous(), torch.empty((B, *oshape, 1), device="cuda"))
    with pytest.raises(luat.LuaError, match="stats must be"):
        cuda.call("CageWarpSmoothL1Volumetric_forwardBackward", ts, tc, tt, empty, empty, gcs, ws, np.zeros((B, 2)), 1.0, ZERO | ONLY_GRID)


def test_affine_sampler_functions(libs):
    """lua/AffineSamplerVolumetric.lua: the grid generator folded into the sampler, on torch.FloatTensor (lua/init.c, host
    tensors staged through the GPU) and torch.CudaTensor (lua/init.cu) -- `output` has the bits of generator + sampler,
    gradInputImages <= 1e-4, gradT <= 2e-5 of the oracle chain."""
    from tests.test_gpu_rowwarp import random_affine
    ours, cuda, _ = libs
    rng = np.random.default_rng(29)
    B, C, ishape, oshape = 5, 2, (7, 8, 9), (6, 9, 34)
    vol = synth.volume("v1", B, *ishape, C, rng)
    T = random_affine(B, rng)
    go = synth.grad_output(B, *oshape, C, rng)
    grid = np.empty((B, *oshape, 4), np.float32)
    assert ours.call("VolumetricGridGenerator_updateOutput", T, grid) == 1          # the generator's own output on this library
    ref_out = oracle.sampler_forward(vol, grid)
    ref_gv, ref_gg = oracle.sampler_backward(vol, grid, go)
    ref_gT = oracle.gridgen_backward(ref_gg)
    # host tensors
    out, gv, gT = np.full((B, *oshape, C), 7, np.float32), np.full_like(vol, 7), np.full_like(T, 7)
    assert ours.call("AffineSamplerVolumetric_updateOutput", vol, T, out) == 1
    assert bits_equal(out, ref_out)
    assert ours.call("AffineSamplerVolumetric_updateGradInput", vol, T, gv, gT, go, _abi.BWD_ZERO_GRADVOL) == 1
    assert max_rel(gv, ref_gv) <= 1e-4 and max_rel(gT, ref_gT) <= 2e-5
    with pytest.raises(luat.LuaError, match="nil value"):                              # float only: not registered for Double
        ours.call("AffineSamplerVolumetric_updateOutput", vol.astype(np.float64), T.astype(np.float64), out.astype(np.float64))
    # CudaTensors; the backward's workspace tensor is module-owned and resized by the function
    tv, tT, tgo = (torch.from_numpy(x).cuda() for x in (vol, T, go))
    tout, tgv, tgT = torch.full((B, *oshape, C), 7.0, device="cuda"), torch.full_like(tv, 7.0), torch.full_like(tT, 7.0)
    ws = luat.Tensor(torch.empty(1 << 18, device="cuda"))
    assert cuda.call("AffineSamplerVolumetric_updateOutput", tv, tT, tout) == 1
    assert bits_equal(tout.cpu().numpy(), ref_out)
    assert cuda.call("AffineSamplerVolumetric_updateGradInput", tv, tT, tgv, tgT, tgo, ws, _abi.BWD_ZERO_GRADVOL) == 1
    assert max_rel(tgv.cpu().numpy(), ref_gv) <= 1e-4 and max_rel(tgT.cpu().numpy(), ref_gT) <= 2e-5
    empty = torch.empty(0, device="cuda")
    tgT2 = torch.full_like(tT, 3.0)
    assert cuda.call("AffineSamplerVolumetric_updateGradInput", tv, tT, empty, tgT2, tgo, ws, _abi.BWD_ZERO_GRADVOL | _abi.BWD_ONLY_GRID) == 1
    assert max_rel(tgT2.cpu().numpy(), ref_gT) <= 2e-5
    with pytest.raises(luat.LuaError, match="affine transform matrices"):
        cuda.call("AffineSamplerVolumetric_updateOutput", tv, tT[:, :3].contiguous(), tout)


def test_sampler_smooth_l1_step_function(libs):
    """lua/BilinearSamplerSmoothL1Volumetric.lua: the one-kernel step with a grid tensor through the fake stack -- forward bits,
    loss, exact IoU counts, gradGrids bit-exact against the oracle fed with the criterion's gradient (G = 4 and 3)."""
    _, cuda, _ = libs
    rng = np.random.default_rng(37)
    B, ishape, oshape = 3, (8, 9, 11), (6, 9, 20)
    vol = synth.volume("v2", B, *ishape, 1, rng)
    tgt = synth.volume("v2", B, *oshape, 1, rng)
    for G in (4, 3):
        grid = synth.grid("g2", B, *oshape, rng, G=G, csz=4)
        ref_out = oracle.sampler_forward(vol, grid, G)
        ref_gv, ref_gg = oracle.sampler_backward(vol, grid, oracle.smooth_l1_backward(ref_out, tgt), G)
        n = int(np.prod(tgt.shape))
        tv, tg, tt = (torch.from_numpy(x).cuda() for x in (vol, grid, tgt))
        kept, gv, gg = torch.full((B, *oshape, 1), 9.0, device="cuda"), torch.full_like(tv, 9.0), torch.full_like(tg, 9.0)
        stats = np.full((B, 3), -1.0)
        ws = luat.Tensor(torch.empty(64, device="cuda"))
        scale = float(np.float32(1.0 / np.float32(n)))
        assert cuda.call("BilinearSamplerSmoothL1Volumetric_forwardBackward", tv, tg, tt, kept, gv, gg, ws, stats, scale,
                         _abi.BWD_ZERO_GRADVOL) == 1
        assert bits_equal(kept.cpu().numpy(), ref_out) and bits_equal(gg.cpu().numpy(), ref_gg), G
        assert max_rel(gv.cpu().numpy(), ref_gv) <= 1e-4
        assert abs(stats[:, 0].sum() / n - oracle.smooth_l1_forward(ref_out, tgt)) <= 1e-9
        assert np.array_equal(stats[:, 1:].astype(np.uint64), oracle.iou_counts(ref_out, tgt))
