"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/volstn_b200.h
declares; the ctypes binding declares exactly the same set.  No compute call is made here."""
import ctypes
import os
import re

import pytest

from alignet_b200 import _abi, build as build_mod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "volstn_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(volstn_b200_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    build_mod.build()
    return _abi.lib()


def test_header_declares_expected_entry_points():
    names = _declared()
    for must in ("volstn_b200_sampler_forward", "volstn_b200_sampler_backward", "volstn_b200_gridgen_forward",
                 "volstn_b200_gridgen_backward", "volstn_b200_cumsum_forward", "volstn_b200_cumsum_backward",
                 "volstn_b200_host_sampler_forward", "volstn_b200_host_sampler_backward"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    raw = ctypes.CDLL(_abi.LIB_PATH)
    for name in _declared():
        assert hasattr(raw, name), f"{name} declared in include/volstn_b200.h but not exported"


def test_binding_matches_header(lib):
    assert sorted(_abi.SIGNATURES) == _declared()


def test_no_torch_in_abi_signatures():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)  # code only, comments cite the reference
    assert "#include <torch" not in src and "ATen" not in src and "THTensor" not in src and "at::" not in src
    assert set(re.findall(r"#include <([^>]+)>", src)) == {"stddef.h", "stdint.h"}


def test_utilities_work_without_gpu(lib):
    assert lib.volstn_b200_abi_version() == 1
    assert lib.volstn_b200_strerror(0) == b"ok"
    assert b"no sm_100" in lib.volstn_b200_strerror(_abi.ERR_NO_DEVICE)
    assert lib.volstn_b200_launch_count() >= 0
    assert lib.volstn_b200_gridgen_backward_workspace(4, 8, 8, 8) > 0


def test_argument_validation_needs_no_gpu(lib):
    """Shape / argument errors are detected before anything touches the device."""
    T5 = _abi.Tensor5

    def desc(shape, data=1 << 20):
        d = T5()
        d.data = data
        st = 1
        for i in range(4, -1, -1):
            d.size[i] = shape[i]
            d.stride[i] = st
            st *= shape[i]
        return d

    vol, grid, out = desc((2, 4, 4, 4, 1)), desc((2, 3, 3, 3, 4)), desc((2, 3, 3, 3, 1))
    assert lib.volstn_b200_sampler_forward(0, 5, vol, grid, out, 0, None) == _abi.ERR_ARG      # bad G
    assert lib.volstn_b200_sampler_forward(7, 4, vol, grid, out, 0, None) == _abi.ERR_ARG      # bad dtype
    bad_grid = desc((3, 3, 3, 3, 4))
    assert lib.volstn_b200_sampler_forward(0, 4, vol, bad_grid, out, 0, None) == _abi.ERR_SHAPE  # lua:33
    g3 = desc((2, 3, 3, 3, 3))
    assert lib.volstn_b200_sampler_forward(0, 4, vol, g3, out, 0, None) == _abi.ERR_SHAPE        # lua:34
    bad_out = desc((2, 3, 3, 2, 1))
    assert lib.volstn_b200_sampler_forward(0, 4, vol, grid, bad_out, 0, None) == _abi.ERR_SHAPE  # lua:73
    # the HOST entry points honour strides too (...c:460-473), but an OUTPUT whose elements alias each other, a
    # negative stride, or a layout too fragmented for DMA copies is refused before the context is touched
    aliased = desc((2, 3, 3, 3, 1))
    aliased.stride[3] = 0
    assert lib.volstn_b200_host_sampler_forward(ctypes.c_void_p(1), 0, 4, vol, grid, aliased, 0) == _abi.ERR_LAYOUT
    neg = desc((2, 3, 3, 3, 4))
    neg.stride[2] = -12
    assert lib.volstn_b200_host_sampler_forward(ctypes.c_void_p(1), 0, 4, vol, neg, out, 0) == _abi.ERR_LAYOUT
    frag_v, frag_g, frag_o = desc((1, 4, 4, 4, 2)), desc((1, 256, 256, 4, 4)), desc((1, 256, 256, 4, 2))
    # three irregular levels: runs of 2 elements, pitched over W, and 65536 (z, y) pairs that each need their own copy
    frag_o.stride[3], frag_o.stride[2], frag_o.stride[1], frag_o.stride[0] = 3, 13, 13 * 256 + 1, (13 * 256 + 1) * 256
    assert lib.volstn_b200_host_sampler_forward(ctypes.c_void_p(1), 0, 4, frag_v, frag_g, frag_o, 0) == _abi.ERR_LAYOUT
    # fused modules: float only, generator extents > 1, bounded cage, exclusive ONLY_ flags
    o5 = desc((2, 3, 3, 3, 1))
    assert lib.volstn_b200_affine_sampler_forward(1, 1 << 20, vol, o5, 0, None) == _abi.ERR_UNSUPPORTED
    assert lib.volstn_b200_affine_sampler_forward(0, 1 << 20, vol, desc((2, 1, 3, 3, 1)), 0, None) == _abi.ERR_SHAPE
    assert lib.volstn_b200_affine_sampler_forward(0, 1 << 20, vol, desc((3, 3, 3, 3, 1)), 0, None) == _abi.ERR_SHAPE
    assert lib.volstn_b200_affine_sampler_backward(0, 1 << 20, vol, vol, 1 << 20, o5, 1 << 20, 8, 0, None) == _abi.ERR_WORKSPACE
    assert lib.volstn_b200_affine_sampler_backward(0, 1 << 20, vol, vol, 1 << 20, o5, 1 << 20, 1 << 30, 3, None) == _abi.ERR_ARG
    assert lib.volstn_b200_affine_sampler_backward_workspace(4, 8, 8, 8) > 0
    cd = (ctypes.c_int64 * 3)(8, 8, 32)
    assert lib.volstn_b200_cage_warp_forward(0, 1 << 20, cd, vol, o5, 0, None) == _abi.ERR_UNSUPPORTED
    cd = (ctypes.c_int64 * 3)(8, 0, 8)
    assert lib.volstn_b200_cage_warp_forward(0, 1 << 20, cd, vol, o5, 0, None) == _abi.ERR_SHAPE
    cd = (ctypes.c_int64 * 3)(8, 8, 8)
    assert lib.volstn_b200_cage_warp_backward(0, 1 << 20, cd, vol, vol, 1 << 20, o5, 3, None) == _abi.ERR_ARG
    # empty batch: nothing to do, no launch, ok
    e_vol, e_grid, e_out = desc((0, 4, 4, 4, 1)), desc((0, 3, 3, 3, 4)), desc((0, 3, 3, 3, 1))
    n0 = lib.volstn_b200_launch_count()
    assert lib.volstn_b200_sampler_forward(0, 4, e_vol, e_grid, e_out, 0, None) == _abi.OK
    assert lib.volstn_b200_launch_count() == n0
    assert lib.volstn_b200_gridgen_forward(0, None, None, 2, 1, 4, 4, None) == _abi.ERR_SHAPE   # lua:6 depth > 1
    assert lib.volstn_b200_sampler_backward(0, 4, vol, grid, vol, grid, out, 3, None) == _abi.ERR_ARG  # both ONLY_ flags


def test_no_device_is_loud(lib):
    """On a box without a B200 the library refuses instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.volstn_b200_device_check(0) == _abi.ERR_NO_DEVICE
    ctx = ctypes.c_void_p()
    assert lib.volstn_b200_host_ctx_create(0, ctypes.byref(ctx)) == _abi.ERR_NO_DEVICE
    from alignet_b200 import nn as vnn
    m = vnn.BilinearSamplerVolumetric()
    with pytest.raises(_abi.VolstnError):
        m.updateOutput([torch.zeros(1, 2, 2, 2, 1), torch.zeros(1, 2, 2, 2, 4)])
