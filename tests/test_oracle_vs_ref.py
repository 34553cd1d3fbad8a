"""Live differential test: oracle/volstn_oracle.c vs the reference compiled unmodified
(oracle/_ref).  Skipped only where oracle/_ref could not be built (no /root/reference and no
prebuilt .so) -- the golden-vector test covers that situation."""
import numpy as np
import pytest

from alignet_b200 import synth
from oracle import oracle
from tests._util import bits_equal, describe_mismatch

pytestmark = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("G", [4, 3])
@pytest.mark.parametrize("gk", ["g1", "g2", "g3", "g4", "g5"])
def test_port_equals_reference(dt, G, gk):
    rng = np.random.default_rng(hash((np.dtype(dt).name, G, gk)) % (2**32))
    for C in (1, 2, 4, 7):
        B, ishape, oshape = 2, (7, 9, 8), (6, 7, 10)
        vol = synth.volume("v1", B, *ishape, C, rng, dt)
        grid = synth.grid(gk, B, *oshape, rng, dt, G=G, csz=4)
        go = synth.grad_output(B, *oshape, C, rng, dt)
        a = oracle.sampler_forward(vol, grid, G, "port")
        b = oracle.sampler_forward(vol, grid, G, "ref")
        assert bits_equal(a, b), describe_mismatch(a, b)
        gv1, gg1 = oracle.sampler_backward(vol, grid, go, G, "port")
        gv2, gg2 = oracle.sampler_backward(vol, grid, go, G, "ref")
        assert bits_equal(gg1, gg2), describe_mismatch(gg1, gg2)
        assert bits_equal(gv1, gv2), describe_mismatch(gv1, gv2)


def test_strided_grid_and_gradout_honoured():
    """The CPU reference honours strides of dims 0-3 of grid / gradOut / out (...c:460-473)."""
    rng = np.random.default_rng(7)
    B, C = 2, 3
    vol = synth.volume("v1", B, 5, 6, 7, C, rng)
    big = synth.grid("g3", B, 4, 5, 12, rng)
    grid = big[:, :, :, ::2, :]  # W-stride 8
    gobig = synth.grad_output(B, 4, 5, 12, C, rng)
    go = gobig[:, :, :, ::2, :]
    a = oracle.sampler_forward(vol, grid, 4, "port")
    b = oracle.sampler_forward(vol, grid, 4, "ref")
    c = oracle.sampler_forward(vol, np.ascontiguousarray(grid), 4, "ref")
    assert bits_equal(a, b) and bits_equal(a, c)
    gv1, gg1 = oracle.sampler_backward(vol, grid, go, 4, "port")
    gv2, gg2 = oracle.sampler_backward(vol, grid, go, 4, "ref")
    assert bits_equal(gv1, gv2) and bits_equal(gg1, gg2)
