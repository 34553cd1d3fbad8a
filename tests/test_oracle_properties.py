"""Known answers derivable from the reference source: the only
'golden' facts the reference itself implies."""
import numpy as np

from alignet_b200 import synth
from oracle import oracle


def test_identity_grid_reproduces_volume():
    rng = np.random.default_rng(0)
    vol = synth.volume("v1", 2, 16, 16, 16, 2, rng)
    grid = synth.grid("g1", 2, 16, 16, 16, rng)
    out = oracle.sampler_forward(vol, grid)
    assert np.max(np.abs(out - vol)) < 1e-5


def test_all_out_of_bounds_is_exact_zero():
    rng = np.random.default_rng(1)
    vol = synth.volume("v1", 1, 6, 6, 6, 3, rng)
    grid = rng.uniform(1.5, 4.0, (1, 5, 5, 5, 4)).astype(np.float32)
    go = synth.grad_output(1, 5, 5, 5, 3, rng)
    out = oracle.sampler_forward(vol, grid)
    gv, gg = oracle.sampler_backward(vol, grid, go)
    assert not out.any() and not gv.any() and not gg.any()


def test_corner_coordinates_hit_corner_voxels():
    rng = np.random.default_rng(2)
    vol = synth.volume("v1", 1, 4, 5, 6, 1, rng)
    grid = np.zeros((1, 1, 1, 2, 4), np.float32)
    grid[0, 0, 0, 0, :3] = -1
    grid[0, 0, 0, 1, :3] = 1
    out = oracle.sampler_forward(vol, grid)
    assert out[0, 0, 0, 0, 0] == vol[0, 0, 0, 0, 0]
    assert out[0, 0, 0, 1, 0] == vol[0, -1, -1, -1, 0]


def test_gradgrid_matches_finite_differences_double():
    """...c:792-822 is the analytic derivative of ...c:566-573 (away from cell boundaries)."""
    rng = np.random.default_rng(3)
    vol = synth.volume("v1", 1, 6, 7, 8, 2, rng, np.float64)
    grid = rng.uniform(-0.9, 0.9, (1, 3, 3, 3, 4))
    # keep sample points away from integer coordinates where the function is only C0
    for ax, n in ((0, 6), (1, 7), (2, 8)):
        c = (grid[..., ax] + 1) * (n - 1) / 2
        frac = c - np.floor(c)
        c = np.floor(c) + np.clip(frac, 0.2, 0.8)
        grid[..., ax] = c * 2 / (n - 1) - 1
    go = synth.grad_output(1, 3, 3, 3, 2, rng, np.float64)
    _, gg = oracle.sampler_backward(vol, grid, go)
    eps = 1e-6
    for ax in range(3):
        gp, gm = grid.copy(), grid.copy()
        gp[..., ax] += eps
        gm[..., ax] -= eps
        fd = ((oracle.sampler_forward(vol, gp) - oracle.sampler_forward(vol, gm)) * go).sum(-1) / (2 * eps)
        assert np.max(np.abs(fd - gg[..., ax])) < 1e-6 * max(1.0, np.max(np.abs(gg)))


def test_gradvol_is_adjoint_of_forward_double():
    rng = np.random.default_rng(4)
    vol = synth.volume("v1", 2, 5, 5, 5, 3, rng, np.float64)
    grid = synth.grid("g4", 2, 4, 4, 4, rng, np.float64)
    go = synth.grad_output(2, 4, 4, 4, 3, rng, np.float64)
    out = oracle.sampler_forward(vol, grid)
    gv, _ = oracle.sampler_backward(vol, grid, go)
    assert abs((out * go).sum() - (gv * vol).sum()) < 1e-9 * max(1.0, abs((out * go).sum()))


def test_cumsum_restatement_known_answers():
    """ALIGNet init (weight 0, bias del, beta -1-del; models/alignet_2d.lua:77-80,118-119) makes the
    cage linspace(-1, 1, csz) along each channel's own axis (Cumsum.lua:17-19)."""
    csz, B = 8, 3
    delta = 2.0 / (csz - 1)
    x = np.full((B, 3, csz, csz, csz), delta, np.float32)
    cage = oracle.cumsum_forward(x) + np.float32(-1 - delta)
    lin = np.linspace(-1, 1, csz)
    assert np.allclose(cage[:, 0], lin[None, :, None, None], atol=1e-6)
    assert np.allclose(cage[:, 1], lin[None, None, :, None], atol=1e-6)
    assert np.allclose(cage[:, 2], lin[None, None, None, :], atol=1e-6)
    # backward is the adjoint (suffix sum)
    rng = np.random.default_rng(5)
    a = rng.standard_normal((B, 3, csz, csz, csz))
    g = rng.standard_normal((B, 3, csz, csz, csz))
    assert abs((oracle.cumsum_forward(a) * g).sum() - (a * oracle.cumsum_backward(g)).sum()) < 1e-9
    # 2-D variant used by the shipped 2-D model (B x 2 x h x w)
    a2 = rng.standard_normal((B, 2, 5, 7))
    o2 = oracle.cumsum_forward(a2)
    assert np.allclose(o2[:, 0], np.cumsum(a2[:, 0], 1)) and np.allclose(o2[:, 1], np.cumsum(a2[:, 1], 2))


def test_gridgen_restatement_known_answers():
    """Identity T gives the base grid (VolumetricGridGenerator.lua:17-20); backward is the adjoint."""
    D, H, W = 5, 6, 7
    T = np.tile(np.eye(4, dtype=np.float32), (2, 1, 1))
    out = oracle.gridgen_forward(T, D, H, W)
    base = oracle.gridgen_base(D, H, W)
    assert np.array_equal(out[0], base) and np.array_equal(out[1], base)
    assert base[0, 0, 0].tolist() == [-1, -1, -1, 1] and base[-1, -1, -1].tolist() == [1, 1, 1, 1]
    rng = np.random.default_rng(6)
    T = rng.standard_normal((2, 4, 4))
    gg = rng.standard_normal((2, D, H, W, 4))
    lhs = (oracle.gridgen_forward(T, D, H, W) * gg).sum()
    rhs = (oracle.gridgen_backward(gg) * T).sum()
    assert abs(lhs - rhs) < 1e-9 * max(1.0, abs(lhs))
    # the 90-degree rotation of stn3dtorch/test.lua:87 (azi = 90) permutes/flips axes of the base grid
    R = np.eye(4)
    R[1, 1], R[1, 2], R[2, 1], R[2, 2] = 0, 1, -1, 0
    g = oracle.gridgen_forward(R[None], 4, 4, 4)
    assert np.allclose(g[0, ..., 1], oracle.gridgen_base(4, 4, 4, np.float64)[..., 2])
    assert np.allclose(g[0, ..., 2], -oracle.gridgen_base(4, 4, 4, np.float64)[..., 1])


# ---------------------------------------------------------------------------------------------
# the restatements behind the fused modules
# ---------------------------------------------------------------------------------------------
def test_smooth_l1_restatement_agrees_with_an_independent_implementation():
    """THNN's SmoothL1Criterion is not vendored (parity unpinned); ATen's CPU smooth_l1_loss implements the same
    published formula independently, so the two must agree (value in double, gradient in float)."""
    import torch
    rng = np.random.default_rng(21)
    x = (3 * rng.standard_normal((2, 5, 6, 7, 1))).astype(np.float32)
    t = (rng.standard_normal((2, 5, 6, 7, 1)) >= 0.5).astype(np.float32)
    for size_average in (True, False):
        # THNN subtracts in `real` (float) and accumulates in double: feed the float difference against zero
        tx = torch.from_numpy((x - t).astype(np.float64)).requires_grad_(True)
        loss = torch.nn.functional.smooth_l1_loss(tx, torch.zeros_like(tx), reduction="mean" if size_average else "sum")
        loss.backward()
        assert abs(oracle.smooth_l1_forward(x, t, size_average) - float(loss.detach())) <= 1e-12 * max(1.0, float(loss.detach()))
        g = oracle.smooth_l1_backward(x, t, size_average)
        assert g.dtype == np.float32
        assert np.max(np.abs(g - tx.grad.numpy())) <= 2e-7 * np.max(np.abs(tx.grad.numpy()))
    # the three branches of the gradient: -norm, norm * x, +norm
    g = oracle.smooth_l1_backward(np.array([-3.0, 0.25, 3.0], np.float32), np.zeros(3, np.float32), False)
    assert list(g) == [-1.0, 0.25, 1.0]


def test_iou_counts_follow_util_lua():
    """util.lua:113-117: thresholds at 0.5 (strict), per-sample intersection and union"""
    pred = np.array([[0.6, 0.5, 0.7, 0.1], [0.0, 0.0, 0.0, 0.0]], np.float32).reshape(2, 1, 1, 4, 1)
    tgt = np.array([[1.0, 1.0, 0.0, 0.0], [1.0, 0.0, 0.0, 0.0]], np.float32).reshape(2, 1, 1, 4, 1)
    c = oracle.iou_counts(pred, tgt)
    assert c.tolist() == [[1, 3], [0, 1]]


def test_alignet_initial_cage_warps_to_identity():
    """models/alignet_2d.lua:77-80,118-119 + Cumsum.lua:17-19: weight 0 / bias del gives the cage linspace(-1, 1, csz),
    whose trilinear up-sample at the identity field is the identity field again: the chain reproduces the source."""
    csz, N = 6, 12
    delta = 2.0 / (csz - 1)
    d = np.full((1, 3, csz, csz, csz), delta, np.float32)
    cage = oracle.cumsum_forward(d) + np.float32(-1 - delta)
    ax = np.linspace(-1, 1, csz)
    assert np.max(np.abs(cage[0, 0] - ax[:, None, None])) < 1e-6 and np.max(np.abs(cage[0, 2] - ax[None, None, :])) < 1e-6
    base = oracle.gridgen_base(N, N, N)[None]
    cage_vol = np.ones((1, csz, csz, csz, 4), np.float32)
    cage_vol[..., :3] = np.moveaxis(cage, 1, -1)
    field = oracle.sampler_forward(cage_vol, np.ascontiguousarray(base))
    assert np.max(np.abs(field[..., :3] - base[..., :3])) < 1e-6
    rng = np.random.default_rng(4)
    src = synth.volume("v1", 1, N, N, N, 1, rng)
    assert np.max(np.abs(oracle.sampler_forward(src, field) - src)) < 2e-5


def test_cumsum_restatement_agrees_with_aten_cpu():
    """Torch7's TH cumsum is not vendored (parity unpinned).  ATen's CPU cumsum is its direct descendant and keeps the
    semantics the restatement assumes (float accumulated in double, each prefix rounded to float): bit-for-bit equal."""
    import torch
    rng = np.random.default_rng(31)
    x = (10 * rng.standard_normal((3, 3, 7, 8, 9))).astype(np.float32)
    ref = oracle.cumsum_forward(x)
    for i in range(3):
        got = torch.cumsum(torch.from_numpy(x[:, i]), dim=1 + i).numpy()
        assert np.array_equal(got.view(np.uint32), ref[:, i].view(np.uint32)), i
    g = rng.standard_normal(x.shape).astype(np.float32)
    refb = oracle.cumsum_backward(g)
    for i in range(3):
        t = torch.from_numpy(g[:, i])
        got = torch.flip(torch.cumsum(torch.flip(t, [1 + i]), dim=1 + i), [1 + i]).numpy()   # Cumsum.lua:44-47
        assert np.array_equal(got.view(np.uint32), refb[:, i].view(np.uint32)), i


def test_compute_iou_reference_grouping_is_modulo_batch():
    """util.lua:115-116 views the B-major tensor as (-1, B) and sums over dim 1: column c = flat indices congruent to c
    modulo B, not sample c.  The restatement reproduces that; the per-sample counts the fused kernels produce are a
    documented deviation.  The torch mirror (alignet_b200.nn.computeIOU) equals the restatement."""
    import torch
    from alignet_b200 import nn as vnn
    rng = np.random.default_rng(12)
    B, n = 3, 8
    pred = rng.random((B, 1, 2, 2, 2)).astype(np.float32)
    tgt = (rng.random((B, 1, 2, 2, 2)) > 0.4).astype(np.float32)
    ref = oracle.iou_reference(pred, tgt)
    # by hand, from the definition
    t, e = (tgt.reshape(-1) > 0.5), (pred.reshape(-1) > 0.5)
    by_hand = 0.0
    for c in range(B):
        idx = np.arange(c, B * n, B)
        by_hand += float((t[idx] & e[idx]).sum()) / float((t[idx] | e[idx]).sum())
    assert abs(ref - by_hand / B) <= 1e-15
    counts = oracle.iou_counts(pred, tgt).astype(np.float64)
    per_sample = float(np.mean(counts[:, 0] / counts[:, 1]))
    assert abs(per_sample - ref) > 1e-6                      # the two groupings are different numbers
    got = float(vnn.computeIOU(torch.from_numpy(pred), torch.from_numpy(tgt)))
    assert abs(got - ref) <= 1e-15


def test_gridgen_restatement_agrees_with_aten_bmm():
    """nn.VolumetricGridGenerator is Lua over torch.bmm / baddbmm (VolumetricGridGenerator.lua:58-60, 79-82) -- Torch7 TH,
    not vendored, parity UNPINNED.  ATen's CPU bmm / baddbmm are TH's direct descendants (same BLAS call underneath): the
    numpy restatement must agree with the Lua module's own sequence of operations executed on them."""
    import torch
    rng = np.random.default_rng(31)
    B, D, H, W = 3, 5, 6, 7
    T = rng.standard_normal((B, 4, 4)).astype(np.float32)
    base = torch.from_numpy(oracle.gridgen_base(D, H, W))                            # lua:12-23
    batch_grid = base.view(1, D, H, W, 4).repeat(B, 1, 1, 1, 1)                       # lua:50-55
    out = torch.bmm(batch_grid.view(B, D * H * W, 4), torch.from_numpy(T).transpose(1, 2))   # lua:58-60
    ref = oracle.gridgen_forward(T, D, H, W)
    assert np.max(np.abs(out.view(B, D, H, W, 4).numpy() - ref)) <= 1e-6
    gg = rng.standard_normal((B, D, H, W, 4)).astype(np.float32)
    grad_t = torch.zeros(B, 4, 4).baddbmm(torch.from_numpy(gg).view(B, D * H * W, 4).transpose(1, 2), batch_grid.view(B, D * H * W, 4))  # lua:79-82
    assert max_rel_np(grad_t.numpy(), oracle.gridgen_backward(gg)) <= 1e-5


def max_rel_np(a, b):
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))) / max(1e-30, float(np.max(np.abs(b)))))
