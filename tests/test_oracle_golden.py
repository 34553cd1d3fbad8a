"""The C restatement (oracle/volstn_oracle.c) against the committed golden vectors, which were
produced by the reference's own C file (tests/golden/make_golden.py).  Bit-exact, float and double,
G = 4 (...c:442-829) and G = 3 (...c:9-439), including NaN / Inf / out-of-int-range coordinates."""
import numpy as np
import pytest

from oracle import oracle
from tests._util import bits_equal, describe_mismatch


def test_golden_has_all_cases(golden):
    names = {k[0] for k in golden}
    assert {"identity_8", "alignet_smooth", "uniform_c1", "uniform_c5", "extreme_c4", "special", "id_64_floor"} <= names
    assert len(golden) == 13 * 2 * 2


def test_port_forward_matches_golden(golden):
    for (name, G, dt), c in golden.items():
        out = oracle.sampler_forward(c["vol"], c["grid"], G, "port")
        assert bits_equal(out, c["out"]), (name, G, dt, describe_mismatch(out, c["out"]))


def test_port_backward_matches_golden(golden):
    for (name, G, dt), c in golden.items():
        gv, gg = oracle.sampler_backward(c["vol"], c["grid"], c["gradOut"], G, "port")
        assert bits_equal(gg, c["gradGrid"]), (name, G, dt, describe_mismatch(gg, c["gradGrid"]))
        assert bits_equal(gv, c["gradVol"]), (name, G, dt, describe_mismatch(gv, c["gradVol"]))


def test_golden_identity_floor_quirk(golden):
    """with an identity grid at N=64, float coordinates land 1 ulp below the integer for
    23 of the 64 positions and floor to k-1.  The index dump must show exactly that."""
    c = golden[("id_64_floor", 4, "float32")]
    idx, mask = oracle.sampler_indices(c["vol"].shape, c["grid"])
    x0 = idx[:64, 2]
    low = int(np.sum(x0 != np.arange(64)))
    assert low == 23, low
    assert np.all((x0 == np.arange(64)) | (x0 == np.arange(64) - 1))


def test_gradgrid_4th_component_zero(golden):
    for (name, G, dt), c in golden.items():
        if G == 4:
            assert np.all(c["gradGrid"][..., 3] == 0)  # ...c:822
