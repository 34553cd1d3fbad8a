"""nn.VolumetricGridGenerator and nn.Cumsum kernels against the numpy restatements (parity
unpinned: Torch7 TH is not available; see oracle/oracle.py).

Tolerances: grid generator output <= 1e-6 absolute (values in [-2, 2]); gradT <= 1e-5 * max|ref|;
Cumsum bit-exact against double-accumulate-then-round (TH CPU semantics)."""
import numpy as np
import pytest
import torch

from alignet_b200 import nn as vnn
from oracle import oracle
from tests._util import bits_equal, max_rel

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("shape", [(5, 6, 7), (32, 32, 32), (2, 2, 2), (3, 40, 70), (64, 64, 64)])
def test_gridgen_forward(shape):
    D, H, W = shape
    rng = np.random.default_rng(D + H + W)
    B = 3
    m = vnn.VolumetricGridGenerator(D, H, W)
    eye = np.tile(np.eye(4, dtype=np.float32), (B, 1, 1))
    out = m.updateOutput(dev(eye)).cpu().numpy()
    base = oracle.gridgen_base(D, H, W)
    for b in range(B):
        assert bits_equal(out[b], base)              # identity T => exactly the Lua base grid (lua:17-20)
    T = rng.standard_normal((B, 4, 4)).astype(np.float32)
    out = m.updateOutput(dev(T)).cpu().numpy()
    assert np.max(np.abs(out - oracle.gridgen_forward(T, D, H, W, np.float64))) <= 1e-6 * max(1.0, np.abs(T).max() * 4)
    assert np.max(np.abs(out - oracle.gridgen_forward(T, D, H, W, np.float32))) <= 1e-6 * max(1.0, np.abs(T).max() * 4)
    # 2-D 4x4 input (lua:39-43)
    o2 = m.updateOutput(dev(T[0])).cpu().numpy()
    assert o2.shape == (D, H, W, 4) and bits_equal(o2, out[0])
    assert bits_equal(m.baseGrid.cpu().numpy(), base)


def test_gridgen_double_and_asserts():
    m = vnn.VolumetricGridGenerator(4, 5, 6)
    T = np.random.default_rng(0).standard_normal((2, 4, 4))
    out = m.updateOutput(dev(T)).cpu().numpy()
    assert np.max(np.abs(out - oracle.gridgen_forward(T, 4, 5, 6))) <= 1e-14 * 10
    with pytest.raises(AssertionError):
        vnn.VolumetricGridGenerator(1, 5, 6)          # lua:6 assert(depth > 1)
    with pytest.raises(AssertionError):
        m.updateOutput(torch.zeros(2, 3, 4, device="cuda"))  # lua:44-47


@pytest.mark.parametrize("shape,B", [((5, 6, 7), 3), ((32, 32, 32), 8), ((64, 64, 64), 2), ((9, 33, 65), 1)])
def test_gridgen_backward(shape, B):
    D, H, W = shape
    rng = np.random.default_rng(B + W)
    m = vnn.VolumetricGridGenerator(D, H, W)
    T = dev(rng.standard_normal((B, 4, 4)).astype(np.float32))
    gg = rng.standard_normal((B, D, H, W, 4)).astype(np.float32)
    gt = m.updateGradInput(T, dev(gg)).cpu().numpy()
    ref = oracle.gridgen_backward(gg, np.float64)
    assert gt.shape == (B, 4, 4)
    assert max_rel(gt, ref) <= 1e-5
    gt2 = m.updateGradInput(T, dev(gg)).cpu().numpy()
    assert bits_equal(gt, gt2)                       # deterministic reduction tree
    if B == 1:
        g1 = m.updateGradInput(T[0], dev(gg)[0])
        assert tuple(g1.shape) == (4, 4)


@pytest.mark.parametrize("shape", [(4, 3, 8, 8, 8), (64, 3, 8, 8, 8), (5, 3, 12, 12, 12), (3, 3, 2, 2, 2),
                                   (2, 3, 5, 7, 11), (3, 2, 9, 10), (2, 2, 40, 70), (2, 1, 100), (2, 3, 4, 5, 33)])
def test_cumsum_forward_backward(shape):
    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal(shape).astype(np.float32)
    g = rng.standard_normal(shape).astype(np.float32)
    m = vnn.Cumsum()
    out = m.updateOutput(dev(x)).cpu().numpy()
    assert bits_equal(out, oracle.cumsum_forward(x, np.float64))
    assert max_rel(out, oracle.cumsum_forward(x, np.float32)) <= 1e-6
    gin = m.updateGradInput(dev(x), dev(g)).cpu().numpy()
    assert bits_equal(gin, oracle.cumsum_backward(g, np.float64))
    # double tensors
    xd = x.astype(np.float64)
    outd = vnn.Cumsum().updateOutput(dev(xd)).cpu().numpy()
    assert max_rel(outd, oracle.cumsum_forward(xd)) <= 1e-14


def test_cumsum_alignet_identity_cage():
    """models/alignet_2d.lua:77-80,102-112: delta = 2/(csz-1) everywhere, + beta = -1-delta => linspace(-1,1)."""
    csz, B = 8, 4
    delta = 2.0 / (csz - 1)
    x = torch.full((B, 3, csz, csz, csz), delta, device="cuda")
    cage = vnn.Cumsum().updateOutput(x) + (-1 - delta)
    lin = torch.linspace(-1, 1, csz, device="cuda")
    assert torch.allclose(cage[:, 0], lin[None, :, None, None].expand(B, csz, csz, csz), atol=1e-6)
    assert torch.allclose(cage[:, 2], lin[None, None, None, :].expand(B, csz, csz, csz), atol=1e-6)


def test_cumsum_reference_quirks():
    m = vnn.Cumsum()
    with pytest.raises(ValueError):
        m.updateOutput(torch.zeros(1, 3, 4, 4, 4, device="cuda"))   # B = 1: squeeze() quirk, Cumsum.lua:18
    ok = vnn.Cumsum(allowSingletonBatch=True).updateOutput(torch.ones(1, 3, 4, 4, 4, device="cuda"))
    assert float(ok[0, 0, 3, 0, 0]) == 4.0 and float(ok[0, 2, 0, 0, 3]) == 4.0
    with pytest.raises(ValueError):
        m.updateOutput(torch.zeros(2, 2, 4, 4, 4, device="cuda"))   # N must equal the number of spatial dims
    # non-contiguous gradOutput is made contiguous first (Cumsum.lua:26-30)
    x = torch.randn(2, 3, 4, 5, 6, device="cuda")
    g = torch.randn(2, 3, 4, 6, 5, device="cuda").transpose(3, 4)
    gin = m.updateGradInput(x, g)
    assert bits_equal(gin.cpu().numpy(), oracle.cumsum_backward(g.contiguous().cpu().numpy()))
