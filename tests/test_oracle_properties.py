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
