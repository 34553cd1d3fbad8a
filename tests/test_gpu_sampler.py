"""Parity of the sm_100a sampler kernels with the reference, through the C ABI (via the nn mirror).

Tolerances (per tensor, SURVEY.md section 8d / BASELINE.json north_star):
  floor indices (z0,y0,x0), 8-bit corner mask : bit-exact
  out, gradGrid                               : bit-exact (stronger than the required 1e-5 relative)
  gradVol (atomic accumulation order)         : max|d| <= 1e-4 * max|ref|  (float), 1e-12 (double)
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

GV_TOL = {np.dtype(np.float32): 1e-4, np.dtype(np.float64): 1e-12}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def cuda_forward(vol, grid, G=4, **kw):
    m = vnn.BilinearSamplerVolumetric(gridWidth=G, **kw)
    return m.updateOutput([dev(vol), dev(grid)]).cpu().numpy()


def cuda_backward(vol, grid, go, G=4, **kw):
    m = vnn.BilinearSamplerVolumetric(gridWidth=G, **kw)
    gv, gg = m.updateGradInput([dev(vol), dev(grid)], dev(go))
    return gv.cpu().numpy(), gg.cpu().numpy()


def assert_gradvol_close(gv, ref, where=""):
    assert np.array_equal(np.isnan(gv), np.isnan(ref)), where
    assert max_rel(gv, ref) <= GV_TOL[ref.dtype], (where, max_rel(gv, ref))


# ---------------------------------------------------------------------------------------------
# golden vectors produced by the reference's own C code
# ---------------------------------------------------------------------------------------------
def test_forward_matches_reference_golden(golden):
    for (name, G, dt), c in golden.items():
        out = cuda_forward(c["vol"], c["grid"], G)
        assert bits_equal(out, c["out"]), (name, G, dt, describe_mismatch(out, c["out"]))


def test_backward_matches_reference_golden(golden):
    for (name, G, dt), c in golden.items():
        gv, gg = cuda_backward(c["vol"], c["grid"], c["gradOut"], G)
        assert bits_equal(gg, c["gradGrid"]), (name, G, dt, describe_mismatch(gg, c["gradGrid"]))
        assert_gradvol_close(gv, c["gradVol"], (name, G, dt))


# ---------------------------------------------------------------------------------------------
# differential vs the oracle on seeded random inputs: every channel path, both grid widths,
# every grid family, ragged sizes, all voxels-per-thread variants
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("G", [4, 3])
@pytest.mark.parametrize("C", [1, 2, 3, 4, 6])
@pytest.mark.parametrize("gk", ["g1", "g2", "g3", "g4", "g5"])
def test_float_parity_random(G, C, gk):
    rng = np.random.default_rng(1000 * G + 10 * C + ord(gk[1]))
    B, ishape, oshape = 3, (9, 11, 13), (7, 10, 37)  # ragged: 2590 voxels/sample, not a multiple of the block
    vol = synth.volume("v1", B, *ishape, C, rng)
    grid = synth.grid(gk, B, *oshape, rng, G=G, csz=4)
    go = synth.grad_output(B, *oshape, C, rng)
    ref_out = oracle.sampler_forward(vol, grid, G)
    ref_gv, ref_gg = oracle.sampler_backward(vol, grid, go, G)
    for vpt in (0, 1, 2, 4):
        for algo in (_abi.ALGO_AUTO, _abi.ALGO_TILE):
            out = cuda_forward(vol, grid, G, vpt=vpt, algo=algo)
            assert bits_equal(out, ref_out), (vpt, algo, describe_mismatch(out, ref_out))
    for vpt in (0, 1, 2):
        for algo in (_abi.ALGO_AUTO, _abi.ALGO_DIRECT, _abi.ALGO_TILE, _abi.ALGO_QUAD):
            gv, gg = cuda_backward(vol, grid, go, G, vpt=vpt, algo=algo)
            assert bits_equal(gg, ref_gg), (vpt, algo, describe_mismatch(gg, ref_gg))
            assert_gradvol_close(gv, ref_gv, (vpt, algo))


@pytest.mark.parametrize("G", [4, 3])
def test_double_parity_random(G):
    rng = np.random.default_rng(77 + G)
    for C in (1, 4, 5):
        vol = synth.volume("v1", 2, 6, 7, 8, C, rng, np.float64)
        grid = synth.grid("g4", 2, 5, 6, 9, rng, np.float64, G=G)
        go = synth.grad_output(2, 5, 6, 9, C, rng, np.float64)
        out = cuda_forward(vol, grid, G)
        ref = oracle.sampler_forward(vol, grid, G)
        assert bits_equal(out, ref), describe_mismatch(out, ref)
        gv, gg = cuda_backward(vol, grid, go, G)
        ref_gv, ref_gg = oracle.sampler_backward(vol, grid, go, G)
        assert bits_equal(gg, ref_gg)
        assert_gradvol_close(gv, ref_gv)


def test_indices_and_masks_bit_exact(golden):
    """Floor/corner index and bounds-mask computation must be bit-exact (north_star)."""
    rng = np.random.default_rng(5)
    cases = [(synth.grid(k, 2, 6, 7, 33, rng, G=4), (2, 9, 8, 64, 1)) for k in ("g1", "g2", "g3", "g4", "g5")]
    cases.append((golden[("special", 4, "float32")]["grid"], (1, 3, 4, 5, 2)))
    cases.append((synth.grid("g1", 1, 64, 64, 64, rng), (1, 64, 64, 64, 1)))    # 23-of-64 floor quirk
    cases.append((synth.grid("g1", 1, 2, 2, 128, rng), (1, 128, 128, 128, 1)))
    for grid, vshape in cases:
        vol = torch.empty(vshape, device="cuda")
        idx, mask = vnn.BilinearSamplerVolumetric().indices(vol, dev(grid))
        ridx, rmask = oracle.sampler_indices(vshape, grid)
        assert np.array_equal(idx.cpu().numpy(), ridx)
        assert np.array_equal(mask.cpu().numpy(), rmask)
    gd = golden[("special", 4, "float64")]["grid"]
    idx, mask = vnn.BilinearSamplerVolumetric().indices(torch.empty((1, 3, 4, 5, 2), device="cuda", dtype=torch.float64), dev(gd))
    ridx, rmask = oracle.sampler_indices((1, 3, 4, 5, 2), gd)
    assert np.array_equal(idx.cpu().numpy(), ridx) and np.array_equal(mask.cpu().numpy(), rmask)


# ---------------------------------------------------------------------------------------------
# layouts: strides on dims 0-3 are honoured like the CPU reference (...c:460-473); 4-D lifting
# ---------------------------------------------------------------------------------------------
def test_strided_tensors_generic_path():
    rng = np.random.default_rng(11)
    B, C = 2, 3
    vol = synth.volume("v1", B, 5, 6, 7, C, rng)
    gbig = synth.grid("g3", B, 4, 5, 12, rng)
    gobig = synth.grad_output(B, 4, 5, 12, C, rng)
    tv, tg, tgo = dev(vol), dev(gbig)[:, :, :, ::2, :], dev(gobig)[:, :, :, ::2, :]
    assert not tg.is_contiguous()
    m = vnn.BilinearSamplerVolumetric()
    out = m.updateOutput([tv, tg]).cpu().numpy()
    ref = oracle.sampler_forward(vol, np.ascontiguousarray(gbig[:, :, :, ::2, :]))
    assert bits_equal(out, ref)
    gv, gg = m.updateGradInput([tv, tg], tgo)
    rgv, rgg = oracle.sampler_backward(vol, np.ascontiguousarray(gbig[:, :, :, ::2]), np.ascontiguousarray(gobig[:, :, :, ::2]))
    assert bits_equal(gg.cpu().numpy(), rgg)
    assert_gradvol_close(gv.cpu().numpy(), rgv)


def test_strided_output_and_grads_direct_abi():
    """out / gradVol / gradGrid with padded strides, straight through the C ABI."""
    rng = np.random.default_rng(12)
    B, C = 2, 2
    vol = synth.volume("v1", B, 4, 5, 6, C, rng)
    grid = synth.grid("g4", B, 3, 4, 5, rng)
    go = synth.grad_output(B, 3, 4, 5, C, rng)
    tv, tg, tgo = dev(vol), dev(grid), dev(go)
    out_big = torch.full((B, 3, 4, 7, C), 9.0, device="cuda")
    out = out_big[:, :, :, :5, :]
    gv_big = torch.zeros((B, 4, 5, 8, C), device="cuda")
    gvv = gv_big[:, :, :, :6, :]
    gg_big = torch.full((B, 3, 4, 6, 4), 9.0, device="cuda")
    ggv = gg_big[:, :, :, :5, :]
    lib = _abi.lib()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    d = vnn._desc
    _abi.check(lib.volstn_b200_sampler_forward(0, 4, d(tv), d(tg), d(out), 0, st), "fwd")
    _abi.check(lib.volstn_b200_sampler_backward(0, 4, d(tv), d(tg), d(gvv), d(ggv), d(tgo), 0, st), "bwd")
    torch.cuda.synchronize()
    assert bits_equal(out.cpu().numpy(), oracle.sampler_forward(vol, grid))
    assert torch.all(out_big[:, :, :, 5:, :] == 9.0)  # padding untouched
    rgv, rgg = oracle.sampler_backward(vol, grid, go)
    assert bits_equal(ggv.cpu().numpy(), rgg)
    assert_gradvol_close(gvv.contiguous().cpu().numpy(), rgv)
    assert torch.all(gg_big[:, :, :, 5:, :] == 9.0) and torch.all(gv_big[:, :, :, 6:, :] == 0)


def test_four_dimensional_inputs_are_lifted():
    """lua:61-67,77-79: a single D x H x W x C sample is viewed as a batch of one and un-viewed."""
    rng = np.random.default_rng(13)
    vol = synth.volume("v1", 1, 5, 6, 7, 2, rng)
    grid = synth.grid("g3", 1, 4, 5, 6, rng)
    go = synth.grad_output(1, 4, 5, 6, 2, rng)
    m = vnn.BilinearSamplerVolumetric()
    out = m.updateOutput([dev(vol)[0], dev(grid)[0]])
    assert out.dim() == 4 and bits_equal(out.cpu().numpy(), oracle.sampler_forward(vol, grid)[0])
    gv, gg = m.updateGradInput([dev(vol)[0], dev(grid)[0]], dev(go)[0])
    assert gv.dim() == 4 and gg.dim() == 4
    assert bits_equal(gg.cpu().numpy(), oracle.sampler_backward(vol, grid, go)[1][0])


def test_module_asserts_like_the_lua_class():
    m = vnn.BilinearSamplerVolumetric()
    v, g = torch.zeros(2, 4, 4, 4, 1, device="cuda"), torch.zeros(2, 3, 3, 3, 4, device="cuda")
    with pytest.raises(AssertionError):
        m.updateOutput([v, g[:1]])                                # lua:33 batch
    with pytest.raises(AssertionError):
        m.updateOutput([v, g[..., :3]])                           # lua:34 grids:size(5)==4
    with pytest.raises(AssertionError):
        m.updateOutput([v.permute(0, 2, 1, 3, 4), g])             # lua:30 contiguous
    with pytest.raises(AssertionError):
        m.updateGradInput([v, g], torch.zeros(2, 3, 3, 2, 1, device="cuda"))  # lua:36-41


# ---------------------------------------------------------------------------------------------
# backward modes and ownership rules
# ---------------------------------------------------------------------------------------------
def test_backward_accumulates_into_gradvol_and_assigns_gradgrid():
    """...c:737 `+=` vs ...c:819 `=`: without VOLSTN_BWD_ZERO_GRADVOL the ABI behaves like the reference."""
    rng = np.random.default_rng(21)
    vol = synth.volume("v1", 2, 5, 5, 5, 1, rng)
    grid = synth.grid("g3", 2, 4, 4, 4, rng)
    go = synth.grad_output(2, 4, 4, 4, 1, rng)
    tv, tg, tgo = dev(vol), dev(grid), dev(go)
    gv = torch.ones_like(tv)
    gg = torch.full_like(tg, 5.0)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    d = vnn._desc
    _abi.check(_abi.lib().volstn_b200_sampler_backward(0, 4, d(tv), d(tg), d(gv), d(gg), d(tgo), 0, st), "bwd")
    rgv, rgg = oracle.sampler_backward(vol, grid, go, grad_vol=np.ones_like(vol))
    assert bits_equal(gg.cpu().numpy(), rgg)
    assert_gradvol_close(gv.cpu().numpy(), rgv)


def test_only_grid_and_only_volume_modes():
    rng = np.random.default_rng(22)
    for C in (1, 4, 5):
        vol = synth.volume("v1", 2, 5, 6, 7, C, rng)
        grid = synth.grid("g4", 2, 4, 5, 6, rng)
        go = synth.grad_output(2, 4, 5, 6, C, rng)
        rgv, rgg = oracle.sampler_backward(vol, grid, go)
        gv, gg = cuda_backward(vol, grid, go, onlyGrid=True)
        assert bits_equal(gg, rgg) and not gv.any()
        _, rgg_og = oracle.sampler_backward(vol, grid, go, only_grid=True)
        assert bits_equal(gg, rgg_og)
        gv, gg = cuda_backward(vol, grid, go, onlyVolume=True)
        assert_gradvol_close(gv, rgv)
        assert not gg.any()


def test_scatter_contention_against_double_oracle():
    """All voxels hit the same 8 corners (constant grid): the float atomics must stay within 1e-4 of the
    DOUBLE oracle, the float CPU sum itself is ~4e-5 away."""
    rng = np.random.default_rng(23)
    B, N, C = 1, 32, 4
    vol = synth.volume("v1", B, N, N, N, C, rng)
    grid = np.zeros((B, N, N, N, 4), np.float32)
    grid[..., :3] = np.array([0.25, -0.5, 0.75], np.float32)
    grid[..., 3] = 1
    go = np.abs(synth.grad_output(B, N, N, N, C, rng))
    ref64, _ = oracle.sampler_backward(vol.astype(np.float64), grid.astype(np.float64), go.astype(np.float64))
    for algo in (_abi.ALGO_DIRECT, _abi.ALGO_TILE, _abi.ALGO_AUTO):
        gv, _ = cuda_backward(vol, grid, go, algo=algo)
        assert max_rel(gv.astype(np.float64), ref64) <= 1e-4, (algo, max_rel(gv.astype(np.float64), ref64))


# ---------------------------------------------------------------------------------------------
# BASELINE.json sizes: size-independent properties + a slice against the oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,B,C", [(32, 32, 1), (64, 16, 1), (128, 2, 4)])
def test_full_size_properties(N, B, C):
    rng = np.random.default_rng(N + C)
    vol_np = synth.volume("v1", B, N, N, N, C, rng)
    vol = dev(vol_np)
    m = vnn.BilinearSamplerVolumetric()
    # (1) identity grid reproduces the volume.  The coordinate ((g+1)*(N-1))/2 carries ~N/2 ulp of
    # rounding, so the bound scales with N (3e-6 of scale at N=64): 4e-7 * N.
    gid = dev(synth.grid("g1", B, N, N, N, rng))
    out = m.updateOutput([vol, gid]).clone()
    assert float((out - vol).abs().max()) <= 4e-7 * N * float(vol.abs().max())
    # (2) linearity in the volume: scaling by a power of two is exact
    out2 = m.updateOutput([vol * 4.0, gid])
    assert torch.equal(out2, out * 4.0)
    # (3) everything out of bounds -> exact zeros in out, gradVol, gradGrid
    goob = torch.empty((B, N, N, N, 4), device="cuda").uniform_(1.5, 3.0)
    go = dev(synth.grad_output(B, N, N, N, C, rng))
    assert not m.updateOutput([vol, goob]).any()
    gv, gg = m.updateGradInput([vol, goob], go)
    assert not gv.any() and not gg.any()
    # (4) adjointness <sample(vol), go> == <vol, gradVol> and gradGrid[...,3] == 0, random grid
    gnp = synth.grid("g4", B, N, N, N, rng)
    g = dev(gnp)
    out = m.updateOutput([vol, g]).clone()
    gv, gg = m.updateGradInput([vol, g], go)
    lhs = float((out.double() * go.double()).sum())
    rhs = float((gv.double() * vol.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))
    assert not gg[..., 3].any()
    # (5) out and gradGrid are deterministic run to run; gradVol within tolerance of itself
    out_b = m.updateOutput([vol, g])
    assert torch.equal(out, out_b)
    gv_a, gg_a = gv.clone(), gg.clone()
    gv_b, gg_b = m.updateGradInput([vol, g], go)
    assert torch.equal(gg_a, gg_b)
    assert float((gv_a - gv_b).abs().max()) <= 1e-4 * float(gv_a.abs().max())
    # (6) first and last sample against the oracle, bit-exact
    for b in (0, B - 1):
        ref = oracle.sampler_forward(vol_np[b:b + 1], gnp[b:b + 1])
        assert bits_equal(out[b:b + 1].cpu().numpy(), ref)
        rgv, rgg = oracle.sampler_backward(vol_np[b:b + 1], gnp[b:b + 1], go[b:b + 1].cpu().numpy())
        assert bits_equal(gg_a[b:b + 1].cpu().numpy(), rgg)
        assert_gradvol_close(gv_a[b:b + 1].cpu().numpy(), rgv)


@pytest.mark.parametrize("gk", ["g1", "g2"])
def test_headline_shape_first_middle_last_sample(gk):
    """64^3 x batch 64, C = 1 -- the shape bench.py's headline is quoted on: the packed kernels then run with
    gridDim.y < B, every CTA walking several samples (batch_rows, sampler.cu).  First, middle and last sample of the
    batch against the reference's C code: out and gradGrid bit for bit, gradVol within 1e-4 of max|ref|."""
    N, B = 64, 64
    rng = np.random.default_rng(64 + len(gk))
    vol_np = synth.volume("v2", B, N, N, N, 1, rng)
    grid_np = synth.grid(gk, B, N, N, N, rng)
    go_np = synth.grad_output(B, N, N, N, 1, rng)
    vol, grid, go = dev(vol_np), dev(grid_np), dev(go_np)
    m = vnn.BilinearSamplerVolumetric()
    out = m.updateOutput([vol, grid])
    gv, gg = m.updateGradInput([vol, grid], go)
    for b in (0, 1, B // 2 - 1, B // 2, B - 2, B - 1):
        sl = slice(b, b + 1)
        assert bits_equal(out[sl].cpu().numpy(), oracle.sampler_forward(vol_np[sl], grid_np[sl])), b
        rgv, rgg = oracle.sampler_backward(vol_np[sl], grid_np[sl], go_np[sl])
        assert bits_equal(gg[sl].cpu().numpy(), rgg), b
        assert_gradvol_close(gv[sl].cpu().numpy(), rgv)


def test_largest_admissible_sample_32bit_offsets():
    """The packed kernels form in-sample offsets in WRAPPING 32-bit arithmetic (sampler.cuh: finish_fast, base32,
    CornerRows); admission is `elements per sample < 2^31` (packed_ok).  One sample of 1024 x 1024 x 2047 voxels =
    2 146 435 072 elements sits 1 048 576 below that limit: the corners of its LAST voxels have offsets just under
    2^31 on every tier (fast: interior; border: exactly on the high faces; exact: straddling them).  Checked against
    the reference's C code (it indexes with `int`, ...c:519-529, valid up to exactly this size)."""
    D, H, W = 1024, 1024, 2047
    assert D * H * W < 2 ** 31 <= D * H * (W + 1)
    rng = np.random.default_rng(2047)
    vol_np = np.zeros((1, D, H, W, 1), np.float32)        # calloc: only the touched slabs become resident
    for z in (0, 1, D - 2, D - 1):
        vol_np[0, z] = rng.standard_normal((H, W, 1)).astype(np.float32)
    vol = torch.zeros((1, D, H, W, 1), device="cuda")
    for z in (0, 1, D - 2, D - 1):
        vol[0, z] = torch.from_numpy(vol_np[0, z]).cuda()
    oD, oH, oW = 4, 8, 64
    grid_np = np.ones((1, oD, oH, oW, 4), np.float32)

    def norm(c, n):      # a coordinate whose un-normalised value is ~c (the kernel recomputes it like the reference)
        return (2.0 * c / (n - 1) - 1.0).astype(np.float32)

    # plane 0: strictly inside the last cell block (fast tier: every corner of every lane in bounds)
    grid_np[0, 0, ..., 0] = norm(rng.uniform(D - 2, D - 1.001, (oH, oW)), D)
    grid_np[0, 0, ..., 1] = norm(rng.uniform(H - 3, H - 1.001, (oH, oW)), H)
    grid_np[0, 0, ..., 2] = norm(rng.uniform(W - 40, W - 1.001, (oH, oW)), W)
    # plane 1: exactly on the high faces (+1 -> size-1: border tier, high corners out of bounds with weight 0)
    grid_np[0, 1, ..., 0] = 1.0
    grid_np[0, 1, ..., 1] = np.where(rng.random((oH, oW)) < 0.5, 1.0, norm(rng.uniform(H - 3, H - 1.001, (oH, oW)), H))
    grid_np[0, 1, ..., 2] = np.where(rng.random((oH, oW)) < 0.5, 1.0, norm(rng.uniform(W - 40, W - 1.001, (oH, oW)), W))
    # plane 2: straddling the far corner (exact tier: some corners out of bounds with non-zero weight)
    grid_np[0, 2, ..., 0] = norm(rng.uniform(D - 2, D + 0.5, (oH, oW)), D)
    grid_np[0, 2, ..., 1] = norm(rng.uniform(H - 2, H + 0.5, (oH, oW)), H)
    grid_np[0, 2, ..., 2] = norm(rng.uniform(W - 2, W + 0.5, (oH, oW)), W)
    # plane 3: the near corner and anywhere
    grid_np[0, 3, ..., :3] = rng.uniform(-1.001, 1.001, (oH, oW, 3)).astype(np.float32)
    grid_np[0, 3, :2, :, :3] = -1.0 + rng.uniform(0, 2e-3, (2, oW, 3)).astype(np.float32)
    go_np = synth.grad_output(1, oD, oH, oW, 1, rng)
    ref = oracle.sampler_forward(vol_np, grid_np)
    gv_ref = np.zeros((1, D, H, W, 1), np.float32)
    _, gg_ref = oracle.sampler_backward(vol_np, grid_np, go_np, grad_vol=gv_ref)
    idx, mask = oracle.sampler_indices(vol_np.shape, grid_np)
    grid, go = dev(grid_np), dev(go_np)
    for vpt in (0, 1):                                   # two and one voxel per thread: different CTA / plane mapping
        m = vnn.BilinearSamplerVolumetric(vpt=vpt)
        out = m.updateOutput([vol, grid])
        assert bits_equal(out.cpu().numpy(), ref), describe_mismatch(out.cpu().numpy(), ref)
        gv, gg = m.updateGradInput([vol, grid], go)
        assert bits_equal(gg.cpu().numpy(), gg_ref)
        # gradVol: compare at every cell any corner touches, and nowhere else may anything be non-zero
        cells = set()
        for (z0, y0, x0), mk in zip(idx.tolist(), mask.tolist()):
            for k in range(8):
                if mk >> k & 1:
                    cells.add(((z0 + (k >> 2 & 1)) * H + y0 + (k >> 1 & 1)) * W + x0 + (k & 1))
        cells = np.array(sorted(cells), np.int64)
        assert cells.max() == D * H * W - 1                   # the very last element of the sample is written
        got = gv.view(-1)[torch.from_numpy(cells).cuda()].cpu().numpy()
        want = gv_ref.reshape(-1)[cells]
        assert np.max(np.abs(got - want)) <= 1e-4 * np.max(np.abs(want))
        assert int(torch.count_nonzero(gv)) == int(np.count_nonzero(want))
    di, dm = vnn.BilinearSamplerVolumetric().indices(vol, grid)
    assert np.array_equal(di.cpu().numpy(), idx) and np.array_equal(dm.cpu().numpy(), mask)


def test_small_volume_backward_wide_dynamic_range():
    """VOLSTN_ALGO_AUTO sends a backward whose volume is much smaller than its output (ALIGNet's cage up-sample) to the
    shared-memory kernel, which accumulates in FIXED POINT scaled to max|gradOut| of each 256-voxel brick (header:
    contributions below 2^-23 of that maximum are rounded away).  The contract stays the one stated for gradVol --
    1e-4 of max|ref| -- also when gradOut spans nine decades; VOLSTN_ALGO_DIRECT keeps the small contributions."""
    rng = np.random.default_rng(99)
    B, N, c = 2, 32, 4
    vol = synth.volume("v1", B, c, c, c, 4, rng)
    grid = synth.grid("g1", B, N, N, N, rng)
    go = (1e-3 * synth.grad_output(B, N, N, N, 4, rng)).astype(np.float32)
    go[0, 3, 5, 7, :] = 1e6                                   # one outlier in one brick
    go[1, N // 2:, :, :, :] *= 1e5
    ref64, _ = oracle.sampler_backward(vol.astype(np.float64), grid.astype(np.float64), go.astype(np.float64))
    for algo in (_abi.ALGO_AUTO, _abi.ALGO_DIRECT):
        gv, gg = cuda_backward(vol, grid, go, algo=algo)
        assert max_rel(gv.astype(np.float64), ref64) <= 1e-4, (algo, max_rel(gv.astype(np.float64), ref64))
        _, rgg = oracle.sampler_backward(vol, grid, go)
        assert bits_equal(gg, rgg)
    # the direct kernel resolves sample 1's small half to fp32 accuracy; the fixed-point one only to its stated step
    gv_d, _ = cuda_backward(vol, grid, go, algo=_abi.ALGO_DIRECT)
    assert max_rel(gv_d[1].astype(np.float64), ref64[1]) <= 1e-5


def test_empty_and_degenerate_shapes():
    m = vnn.BilinearSamplerVolumetric()
    v = torch.zeros(0, 4, 4, 4, 2, device="cuda")
    g = torch.zeros(0, 3, 3, 3, 4, device="cuda")
    assert tuple(m.updateOutput([v, g]).shape) == (0, 3, 3, 3, 2)
    gv, gg = m.updateGradInput([v, g], torch.zeros(0, 3, 3, 3, 2, device="cuda"))
    assert tuple(gv.shape) == (0, 4, 4, 4, 2) and tuple(gg.shape) == (0, 3, 3, 3, 4)
    # zero-extent grid: gradVol must still come back zero-filled
    v = torch.ones(2, 4, 4, 4, 1, device="cuda")
    g = torch.zeros(2, 0, 3, 3, 4, device="cuda")
    gv, gg = m.updateGradInput([v, g], torch.zeros(2, 0, 3, 3, 1, device="cuda"))
    assert not gv.any()
    # single-voxel volume: (size-1) == 0 on every axis
    rng = np.random.default_rng(3)
    vol = synth.volume("v1", 2, 1, 1, 1, 3, rng)
    grid = synth.grid("g3", 2, 3, 3, 3, rng)
    assert bits_equal(cuda_forward(vol, grid), oracle.sampler_forward(vol, grid))


def test_launch_counter_counts_our_kernels():
    n0 = _abi.launch_count()
    rng = np.random.default_rng(1)
    cuda_forward(synth.volume("v1", 1, 4, 4, 4, 1, rng), synth.grid("g3", 1, 4, 4, 4, rng))
    assert _abi.launch_count() == n0 + 1


# ---------------------------------------------------------------------------------------------
# the warp-voted tiers of the packed kernels (fast / border / exact) compute the same bits
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C", [1, 4])
def test_tiers_are_bit_identical(C):
    """Grids crafted so that whole warps land in each tier: strictly interior coordinates (fast), coordinates
    exactly on the high faces (border), outside / non-finite coordinates (exact).  Every tier, every VPT and the
    debug flags that force the slower tiers must give the oracle's bits."""
    rng = np.random.default_rng(77)
    B, D, H, W = 2, 6, 8, 64                                     # W = 64: two full warps per row
    vol = synth.volume("v1", B, D, H, W, C, rng)
    go = synth.grad_output(B, D, H, W, C, rng)
    base = synth.grid("g1", B, D, H, W, rng)                     # identity: its last column sits ON the high face
    grids = {
        "interior": (base * np.float32(0.9)).astype(np.float32),                     # fast tier everywhere
        "faces": base,                                                                # border tier for half the warps
        "outside": (base * np.float32(1.3) + np.float32(0.05)).astype(np.float32),    # exact tier near the faces
    }
    nf = base.copy()
    nf[0, 2, 3, 10, 0] = np.nan
    nf[1, 1, 1, 40, 2] = np.inf
    nf[1, 4, 6, 63, 1] = -np.inf
    nf[0, 0, 0, 5, 2] = 3e9                                                           # beyond int32 after scaling
    grids["nonfinite"] = nf
    NO_FAST, NO_BORDER = 0x10000, 0x20000
    for name, grid in grids.items():
        ref_out = oracle.sampler_forward(vol, grid)
        rgv, rgg = oracle.sampler_backward(vol, grid, go)
        for vpt in (1, 2, 4):
            for dbg in (0, NO_BORDER, NO_FAST):
                m = vnn.BilinearSamplerVolumetric(vpt=vpt)
                m._flags = (lambda v, d: (lambda: _abi.ALGO_DIRECT | _abi.VPT(v) | d))(vpt, dbg)
                out = m.updateOutput([dev(vol), dev(grid)]).cpu().numpy()
                assert bits_equal(out, ref_out), (name, vpt, hex(dbg), describe_mismatch(out, ref_out))
                if vpt <= 2:
                    gv, gg = m.updateGradInput([dev(vol), dev(grid)], dev(go))
                    assert bits_equal(gg.cpu().numpy(), rgg), (name, vpt, hex(dbg), describe_mismatch(gg.cpu().numpy(), rgg))
                    assert_gradvol_close(gv.cpu().numpy(), rgv, (name, vpt, hex(dbg)))


@pytest.mark.parametrize("C", [1, 2, 4])
@pytest.mark.parametrize("gk", ["g1", "g2", "g3", "g4"])
def test_small_volume_under_large_output_uses_privatised_backward(C, gk):
    """ALIGNet's cage up-sample in the reference's own module graph (models/alignet_2d.lua:136-147): a csz^3 volume
    sampled at a full-resolution field.  The automatic algorithm is the privatised, warp-aggregated kernel there
    (output voxels >= 128 x volume cells); same contract as everywhere: gradGrid bit-exact, gradVol within 1e-4."""
    rng = np.random.default_rng(300 + C + ord(gk[1]))
    B, ishape, oshape = 3, (3, 5, 4), (20, 18, 40)
    vol = synth.volume("v1", B, *ishape, C, rng)
    grid = synth.grid(gk, B, *oshape, rng, csz=4)
    go = synth.grad_output(B, *oshape, C, rng)
    ref_gv, ref_gg = oracle.sampler_backward(vol, grid, go)
    for kw in (dict(), dict(onlyVolume=True), dict(algo=_abi.ALGO_DIRECT)):
        gv, gg = cuda_backward(vol, grid, go, **kw)
        assert_gradvol_close(gv, ref_gv, (kw, max_rel(gv, ref_gv)))
        if "onlyVolume" not in kw:
            assert bits_equal(gg, ref_gg), (kw, describe_mismatch(gg, ref_gg))
    # a constant gradOut puts every lane of a warp on the same few cells with equal signs: the aggregation's worst case
    ones = np.ones_like(go)
    gv, _ = cuda_backward(vol, grid, ones)
    assert_gradvol_close(gv, oracle.sampler_backward(vol, grid, ones)[0])


# ---------------------------------------------------------------------------------------------
# VOLSTN_ALGO_TMA: the volume box of an 8^3 output brick staged by one TMA load (zero fill = the reference's padding),
# gradVol privatised and flushed by one TMA reduction.  Same contract as the direct kernels.
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("G", [4, 3])
@pytest.mark.parametrize("gk", ["g1", "g2", "g3", "g4", "g5"])
@pytest.mark.parametrize("ishape,oshape", [((9, 11, 12), (7, 10, 37)), ((16, 16, 16), (16, 16, 16)), ((5, 6, 8), (20, 9, 8)), ((32, 32, 32), (24, 24, 24))])
def test_tma_staged_kernels_match_the_reference(G, gk, ishape, oshape):
    rng = np.random.default_rng(600 + G + ord(gk[1]) + ishape[0] + oshape[2])
    B = 3
    vol = synth.volume("v1", B, *ishape, 1, rng)
    grid = synth.grid(gk, B, *oshape, rng, G=G, csz=4)
    if gk == "g4":
        grid[0, 0, 0, :4, :3] = [np.nan, np.inf, -np.inf]           # not stageable: that brick takes the direct code
        grid[1, 1, 1, 1, :3] = 1e30
    go = synth.grad_output(B, *oshape, 1, rng)
    ref_out = oracle.sampler_forward(vol, grid, G)
    ref_gv, ref_gg = oracle.sampler_backward(vol, grid, go, G)
    out = cuda_forward(vol, grid, G, algo=_abi.ALGO_TMA)
    assert bits_equal(out, ref_out), describe_mismatch(out, ref_out)
    gv, gg = cuda_backward(vol, grid, go, G, algo=_abi.ALGO_TMA)
    assert bits_equal(gg, ref_gg), describe_mismatch(gg, ref_gg)
    assert_gradvol_close(gv, ref_gv, (G, gk))
    gv2, gg2 = cuda_backward(vol, grid, go, G, algo=_abi.ALGO_TMA, onlyGrid=True)
    assert bits_equal(gg2, ref_gg) and not gv2.any()


def test_tma_staged_kernels_accumulate_and_use_tma():
    """gradVol is accumulated into (no zero-fill flag) by the TMA reduction like by the direct reductions; the library
    carries the TMA instructions (UTMALDG / UTMAREDG in its SASS is checked on the build side, profiles/r02_sass.md)."""
    rng = np.random.default_rng(611)
    B, N = 2, 24
    vol = synth.volume("v1", B, N, N, N, 1, rng)
    grid = synth.grid("g2", B, N, N, N, rng, csz=4)
    go = synth.grad_output(B, N, N, N, 1, rng)
    init = rng.standard_normal(vol.shape).astype(np.float32)
    rgv, rgg = oracle.sampler_backward(vol, grid, go, grad_vol=init.copy())
    tv, tg, tgo = dev(vol), dev(grid), dev(go)
    gv, gg = dev(init), torch.empty_like(tg)
    d = vnn._desc
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    n0 = _abi.launch_count()
    _abi.check(_abi.lib().volstn_b200_sampler_backward(0, 4, d(tv), d(tg), d(gv), d(gg), d(tgo), _abi.ALGO_TMA, st), "bwd")
    assert _abi.launch_count() - n0 == 1
    assert bits_equal(gg.cpu().numpy(), rgg)
    assert_gradvol_close(gv.cpu().numpy(), rgv)
    # unsupported shapes (W not a multiple of 4, C > 1) quietly take the automatic path
    vol2 = synth.volume("v1", B, 7, 9, 13, 2, rng)
    grid2 = synth.grid("g3", B, 6, 7, 9, rng)
    assert bits_equal(cuda_forward(vol2, grid2, algo=_abi.ALGO_TMA), oracle.sampler_forward(vol2, grid2))
