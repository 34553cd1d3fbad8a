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


# ---------------------------------------------------------------------------------------------
# the fused modules' golden vectors (tests/golden/make_cage_golden.py): chains of the reference's own C code
# ---------------------------------------------------------------------------------------------
def load_cage_golden():
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cage_warp_golden.npz"))
    cases = {}
    for k in z.files:
        name, what = k.split("|")
        cases.setdefault(name, {})[what] = z[k]
    return cases


def test_port_chain_matches_the_reference_chain():
    """the oracle chain used by tests/test_gpu_rowwarp.py (port, twice) reproduces the reference chain bit for bit"""
    cases = load_cage_golden()
    assert {"alignet_8", "alignet_4", "identity_6", "folded", "wide_row"} <= set(cases)
    for name, c in cases.items():
        src, cage, go = c["source"], c["cage"], c["gradOut"]
        B, oshape = src.shape[0], c["out"].shape[1:4]
        base = oracle.gridgen_base(*oshape)
        field_i = np.ascontiguousarray(np.broadcast_to(base, (B,) + base.shape))
        cage_vol = np.ones(cage.shape[:1] + cage.shape[2:] + (4,), np.float32)
        cage_vol[..., :3] = np.moveaxis(cage, 1, -1)
        field = oracle.sampler_forward(cage_vol, field_i)
        assert bits_equal(field, c["field"]), name
        out = oracle.sampler_forward(src, field)
        assert bits_equal(out, c["out"]), name
        gsrc, gfield = oracle.sampler_backward(src, field, go)
        gcv, _ = oracle.sampler_backward(cage_vol, field_i, gfield)
        assert bits_equal(gsrc, c["gradSource"]), name
        assert bits_equal(np.ascontiguousarray(np.moveaxis(gcv[..., :3], -1, 1)), c["gradCage"]), name
        if "step_loss" in c:
            assert abs(oracle.smooth_l1_forward(out, c["target"]) - float(c["step_loss"])) == 0.0
            assert np.array_equal(oracle.iou_counts(out, c["target"]), c["step_iou"])


def test_identity_cage_golden_is_the_identity_warp():
    c = load_cage_golden()["identity_6"]
    assert np.max(np.abs(c["out"] - c["source"])) <= 2e-5
