"""TEST INFRASTRUCTURE: Python driver of tests/luat_harness/harness.c.

``Harness(path)`` loads one glue library built against the stand-in Torch7 headers -- ``lua/build/libvolstn.so``
(lua/init.c, the drop-in), ``lua/build/libcuvolstn.so`` (lua/init.cu) or ``oracle/_ref/libvolstn_refinit.so`` (the
REFERENCE's own stn3dtorch/init.c, compiled unmodified) -- runs its ``luaopen_*`` and calls the registered
``lua_CFunction``s with tensors in the reference's stack slots (self = slot 1; generic/...c:444-446, 587-591).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
DROPIN_CPU = os.path.join(ROOT, "lua", "build", "libvolstn.so")
DROPIN_CUDA = os.path.join(ROOT, "lua", "build", "libcuvolstn.so")
REFERENCE_CPU = os.path.join(ROOT, "oracle", "_ref", "libvolstn_refinit.so")

# generic/BilinearSamplerVolumetric.c:831-836 (and BilinearSamplerVolumetric.cu:1074-1080)
REFERENCE_NAMES = ["BilinearSamplerVolumetric_updateOutput", "BilinearSamplerVolumetric_updateGradInput",
                   "BilinearSamplerBHWD_updateOutput", "BilinearSamplerBHWD_updateGradInput"]
NEW_NAMES = ["Cumsum_updateOutput", "Cumsum_updateGradInput",
             "VolumetricGridGenerator_updateOutput", "VolumetricGridGenerator_updateGradInput"]
# the fused modules of ALIGNet's warp head (lua/CageWarpVolumetric.lua, lua/CageWarpSmoothL1Volumetric.lua): CudaTensor only
FUSED_NAMES = ["CageWarpVolumetric_updateOutput", "CageWarpVolumetric_updateGradInput",
               "CageWarpSmoothL1Volumetric_forwardBackward", "BilinearSamplerSmoothL1Volumetric_forwardBackward"]
# lua/AffineSamplerVolumetric.lua: torch.FloatTensor (host, staged through the GPU) and torch.CudaTensor
AFFINE_NAMES = ["AffineSamplerVolumetric_updateOutput", "AffineSamplerVolumetric_updateGradInput"]


class _THTensor(ctypes.Structure):  # layout of HARNESS_DECLARE_TENSOR in include/TH.h
    _fields_ = [("size", ctypes.POINTER(ctypes.c_long)), ("stride", ctypes.POINTER(ctypes.c_long)),
                ("nDimension", ctypes.c_int), ("data", ctypes.c_void_p), ("capacity", ctypes.c_long)]


class Tensor:
    """A stand-in THTensor over a numpy array (host) or a torch CUDA tensor (device); keeps everything alive."""

    def __init__(self, a, tname=None):
        import torch
        self.obj = a
        if isinstance(a, np.ndarray):
            shape, strides, ptr = a.shape, [s // a.itemsize for s in a.strides], a.ctypes.data
            self.tname = tname or {np.dtype(np.float32): "torch.FloatTensor", np.dtype(np.float64): "torch.DoubleTensor"}[a.dtype]
            cap = a.size
        else:
            assert isinstance(a, torch.Tensor)
            shape, strides, ptr = tuple(a.shape), list(a.stride()), a.data_ptr()
            if a.is_cuda:
                assert a.dtype == torch.float32
                self.tname = tname or "torch.CudaTensor"
            else:
                self.tname = tname or {torch.float32: "torch.FloatTensor", torch.float64: "torch.DoubleTensor"}[a.dtype]
            cap = a.numel()
        n = max(len(shape), 1)
        self._size = (ctypes.c_long * n)(*shape)
        self._stride = (ctypes.c_long * n)(*strides)
        self.th = _THTensor(self._size, self._stride, len(shape), ptr, cap)
        self.tname_b = self.tname.encode()   # the harness keeps this pointer for the duration of the call


class LuaError(RuntimeError):
    pass


class Harness:
    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = l = ctypes.CDLL(path)
        l.harness_new.restype = ctypes.c_void_p
        for f in ("harness_free", "harness_open", "harness_nreg", "harness_stack_height"):
            getattr(l, f).argtypes = [ctypes.c_void_p]
        for f in ("harness_reg_type", "harness_reg_table", "harness_reg_name"):
            getattr(l, f).argtypes = [ctypes.c_void_p, ctypes.c_int]
            getattr(l, f).restype = ctypes.c_char_p
        for f in ("harness_error", "harness_global_set"):
            getattr(l, f).argtypes = [ctypes.c_void_p]
            getattr(l, f).restype = ctypes.c_char_p
        l.harness_args_clear.argtypes = [ctypes.c_void_p, ctypes.c_int]
        l.harness_arg_tensor.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p]
        l.harness_arg_number.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double]
        l.harness_set_stream.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        l.harness_call.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
        self.L = l.harness_new()
        self.opened = l.harness_open(self.L)   # require 'libvolstn' / 'libcuvolstn'
        if self.opened < 0:
            raise LuaError(l.harness_error(self.L).decode())
        self.stack_after_open = l.harness_stack_height(self.L)

    def close(self):
        if self.L:
            self.lib.harness_free(self.L)
            self.L = None

    def global_table(self) -> str:
        return self.lib.harness_global_set(self.L).decode()

    def registrations(self):
        """[(luaT type, table, function name)] in registration order"""
        return [(self.lib.harness_reg_type(self.L, i).decode(), self.lib.harness_reg_table(self.L, i).decode(),
                 self.lib.harness_reg_name(self.L, i).decode()) for i in range(self.lib.harness_nreg(self.L))]

    def call(self, name: str, *args, stream=None):
        """``args[0].nn.<name>(self, *args)``: tensors (numpy / torch) or numbers go to slots 2, 3, ... ; dispatch on the
        luaT type of the first tensor, as the Lua files do (BilinearSamplerVolumetric.lua:74,109)."""
        ts = [a if isinstance(a, (Tensor, int, float, type(None))) else Tensor(a) for a in args]
        first = next(t for t in ts if isinstance(t, Tensor))
        self.lib.harness_args_clear(self.L, 1 + len(ts))
        keep = []
        for i, t in enumerate(ts):
            if isinstance(t, Tensor):
                keep.append(t)
                self.lib.harness_arg_tensor(self.L, 2 + i, t.tname_b, ctypes.addressof(t.th))
            elif t is not None:
                self.lib.harness_arg_number(self.L, 2 + i, float(t))
        self.lib.harness_set_stream(self.L, ctypes.c_void_p(stream or 0))
        rc = self.lib.harness_call(self.L, first.tname_b, name.encode())
        if rc == -2:
            raise LuaError(f"attempt to call field '{name}' (a nil value) of {first.tname}.nn")
        if rc == -1:
            raise LuaError(self.lib.harness_error(self.L).decode())
        return rc
