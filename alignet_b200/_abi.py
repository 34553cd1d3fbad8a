"""ctypes binding of include/volstn_b200.h -- nothing but declarations and error translation.

The library is loaded from alignet_b200/csrc/libvolstn_b200.so (built in-tree by
``python -m alignet_b200.build``).  If it is missing this module raises: there is no fallback
implementation anywhere in the package.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VOLSTN_B200_LIB") or os.path.join(HERE, "csrc", "libvolstn_b200.so")

# status codes / enums / flags (mirror include/volstn_b200.h)
OK, ERR_ARG, ERR_SHAPE, ERR_LAYOUT, ERR_CUDA, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_WORKSPACE = range(8)
F32, F64 = 0, 1
BWD_ONLY_GRID = 0x1
BWD_ONLY_VOLUME = 0x2
BWD_ZERO_GRADVOL = 0x4
HOST_KEEP_INPUTS = 0x8
HOST_REUSE_INPUTS = 0x10
HOST_ASYNC = 0x20
FAST_ARITH = 0x40000
ALGO_AUTO, ALGO_DIRECT, ALGO_TILE, ALGO_QUAD, ALGO_ROW, ALGO_TMA = 0x000, 0x100, 0x200, 0x400, 0x500, 0x600
VPT_SHIFT = 12


def VPT(n: int) -> int:
    return (n & 0xF) << VPT_SHIFT


class Tensor5(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("size", ctypes.c_int64 * 5), ("stride", ctypes.c_int64 * 5)]


class VolstnError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        super().__init__(f"{where}: status {status}: {detail}")


_P5 = ctypes.POINTER(Tensor5)
_vp, _i, _u, _i64, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_int64, ctypes.c_size_t
_pi64 = ctypes.POINTER(ctypes.c_int64)

# name -> (restype, argtypes); every symbol declared in include/volstn_b200.h
SIGNATURES = {
    "volstn_b200_sampler_forward": (_i, [_i, _i, _P5, _P5, _P5, _u, _vp]),
    "volstn_b200_sampler_backward": (_i, [_i, _i, _P5, _P5, _P5, _P5, _P5, _u, _vp]),
    "volstn_b200_sampler_indices": (_i, [_i, _i, _pi64, _P5, _vp, _vp, _vp]),
    "volstn_b200_affine_sampler_forward": (_i, [_i, _vp, _P5, _P5, _u, _vp]),
    "volstn_b200_affine_sampler_backward": (_i, [_i, _vp, _P5, _P5, _vp, _P5, _vp, _sz, _u, _vp]),
    "volstn_b200_affine_sampler_backward_workspace": (_sz, [_i64, _i64, _i64, _i64]),
    "volstn_b200_cage_warp_forward": (_i, [_i, _vp, _pi64, _P5, _P5, _u, _vp]),
    "volstn_b200_cage_warp_backward": (_i, [_i, _vp, _pi64, _P5, _P5, _vp, _P5, _u, _vp]),
    "volstn_b200_cage_warp_step": (_i, [_i, _vp, _pi64, _P5, _P5, _P5, _P5, _vp, _vp, _vp, ctypes.c_float, _u, _vp]),
    "volstn_b200_sampler_step": (_i, [_i, _i, _P5, _P5, _P5, _P5, _P5, _P5, _vp, _vp, ctypes.c_float, _u, _vp]),
    "volstn_b200_gridgen_forward": (_i, [_i, _vp, _vp, _i64, _i64, _i64, _i64, _vp]),
    "volstn_b200_gridgen_backward": (_i, [_i, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _sz, _vp]),
    "volstn_b200_gridgen_backward_workspace": (_sz, [_i64, _i64, _i64, _i64]),
    "volstn_b200_cumsum_forward": (_i, [_i, _vp, _vp, _i64, _i, _pi64, _vp]),
    "volstn_b200_cumsum_backward": (_i, [_i, _vp, _vp, _i64, _i, _pi64, _vp]),
    "volstn_b200_host_ctx_create": (_i, [_i, ctypes.POINTER(_vp)]),
    "volstn_b200_host_ctx_destroy": (_i, [_vp]),
    "volstn_b200_host_pin": (_i, [_vp, _sz]),
    "volstn_b200_host_unpin": (_i, [_vp]),
    "volstn_b200_host_ticket": (ctypes.c_uint64, [_vp]),
    "volstn_b200_host_wait": (_i, [_vp, ctypes.c_uint64]),
    "volstn_b200_host_numa_node": (_i, [_i]),
    "volstn_b200_host_bind_thread": (_i, [_i]),
    "volstn_b200_host_sampler_forward": (_i, [_vp, _i, _i, _P5, _P5, _P5, _u]),
    "volstn_b200_host_sampler_backward": (_i, [_vp, _i, _i, _P5, _P5, _P5, _P5, _P5, _u]),
    "volstn_b200_host_gridgen_forward": (_i, [_vp, _i, _vp, _vp, _i64, _i64, _i64, _i64]),
    "volstn_b200_host_gridgen_backward": (_i, [_vp, _i, _vp, _vp, _i64, _i64, _i64, _i64]),
    "volstn_b200_host_affine_sampler_forward": (_i, [_vp, _i, _vp, _P5, _P5, _u]),
    "volstn_b200_host_affine_sampler_backward": (_i, [_vp, _i, _vp, _P5, _P5, _vp, _P5, _u]),
    "volstn_b200_host_cage_warp_forward": (_i, [_vp, _i, _vp, _pi64, _P5, _P5, _u]),
    "volstn_b200_host_cage_warp_backward": (_i, [_vp, _i, _vp, _pi64, _P5, _P5, _vp, _P5, _u]),
    "volstn_b200_host_cumsum_forward": (_i, [_vp, _i, _vp, _vp, _i64, _i, _pi64]),
    "volstn_b200_host_cumsum_backward": (_i, [_vp, _i, _vp, _vp, _i64, _i, _pi64]),
    "volstn_b200_abi_version": (_i, []),
    "volstn_b200_strerror": (ctypes.c_char_p, [_i]),
    "volstn_b200_last_cuda_error": (_i, []),
    "volstn_b200_last_cuda_error_string": (ctypes.c_char_p, []),
    "volstn_b200_device_check": (_i, [_i]),
    "volstn_b200_launch_count": (ctypes.c_uint64, []),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load the C-ABI library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m alignet_b200.build`. "
                "alignet_b200 has no fallback implementation.")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError here means header and library disagree
            fn.restype = res
            fn.argtypes = args
        if l.volstn_b200_abi_version() != 1:
            raise ImportError("libvolstn_b200.so ABI version mismatch")
        _lib = l
    return _lib


def check(status: int, where: str) -> None:
    if status == OK:
        return
    l = lib()
    msg = l.volstn_b200_strerror(status).decode()
    if status == ERR_CUDA:
        msg += " -- " + l.volstn_b200_last_cuda_error_string().decode()
    raise VolstnError(status, where, msg)


def launch_count() -> int:
    return int(lib().volstn_b200_launch_count())
