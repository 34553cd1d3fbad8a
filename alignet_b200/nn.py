"""Host-side mirror of the reference's Lua ``nn`` modules for the volumetric warp path.

Same class names, constructor arguments, ``updateOutput`` / ``updateGradInput`` protocol, tensor
layouts ([-1, 1] coordinates, channels-last ``B x D x H x W x C`` volumes, ``B x D x H x W x 4``
``(z, y, x, .)`` grids), assertions and field names as

* ``nn.BilinearSamplerVolumetric``  -- stn3dtorch/BilinearSamplerVolumetric.lua
* ``nn.VolumetricGridGenerator``    -- stn3dtorch/VolumetricGridGenerator.lua
* ``nn.Cumsum``                     -- Cumsum.lua

The Lua files dispatch on the tensor type into a native library
(``inputImages.nn.BilinearSamplerVolumetric_updateOutput(self, ...)``, lua:74); these classes
dispatch the same way into the C ABI of include/volstn_b200.h:

* ``torch.cuda.FloatTensor`` / ``DoubleTensor`` -> ``volstn_b200_*``      (device pointers, caller's stream)
* ``torch.FloatTensor`` / ``DoubleTensor`` (host) -> ``volstn_b200_host_*`` (staged through the GPU)

PyTorch is used for memory and streams only.  There is no CPU or eager fallback: without the
built library, or without a B200, every call raises.
"""
from __future__ import annotations

import ctypes
import threading
from typing import List, Optional, Sequence

import torch

from . import _abi
from ._abi import Tensor5

__all__ = ["Module", "host_wait", "host_ticket", "BilinearSamplerVolumetric", "VolumetricGridGenerator", "Cumsum",
           "AffineSamplerVolumetric", "CageWarpVolumetric", "CageWarpSmoothL1Volumetric",
           "BilinearSamplerSmoothL1Volumetric", "computeIOU"]


# ------------------------------------------------------------------------------------------------
# plumbing
# ------------------------------------------------------------------------------------------------
def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _abi.F32
    if t.dtype == torch.float64:
        return _abi.F64
    raise TypeError(f"volstn: unsupported tensor type {t.dtype} (reference registers Float and Double only, init.c:8-19)")


def _desc(t: torch.Tensor) -> Tensor5:
    assert t.dim() == 5
    d = Tensor5()
    d.data = t.data_ptr()
    for i in range(5):
        d.size[i] = t.size(i)
        d.stride[i] = t.stride(i)
    return d


def _stream(t: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


_tls = threading.local()
_checked_devices = set()


def _require_device(index: int) -> None:
    if index in _checked_devices:
        return
    _abi.check(_abi.lib().volstn_b200_device_check(index), f"device {index}")
    _checked_devices.add(index)


def _host_ctx(device: Optional[int] = None) -> ctypes.c_void_p:
    """One staging context per (thread, device) for host tensors."""
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    cache = getattr(_tls, "ctx", None)
    if cache is None:
        cache = _tls.ctx = {}
    if device not in cache:
        ctx = ctypes.c_void_p()
        _abi.check(_abi.lib().volstn_b200_host_ctx_create(device, ctypes.byref(ctx)), "host_ctx_create")
        cache[device] = ctx
    return cache[device]


def host_ticket() -> int:
    """Ticket of the most recent asynchronous host call of this thread (``hostAsync=True`` modules)."""
    return int(_abi.lib().volstn_b200_host_ticket(_host_ctx()))


def host_wait(ticket: int = 0) -> None:
    """Block until the asynchronous host call with that ticket (0: every call) of this thread is complete."""
    _abi.check(_abi.lib().volstn_b200_host_wait(_host_ctx(), ticket), "host_wait")


def _same_place(*ts: torch.Tensor) -> None:
    dev, dt = ts[0].device, ts[0].dtype
    for t in ts[1:]:
        if t.device != dev or t.dtype != dt:
            raise TypeError("volstn: all tensors of a call must share device and dtype")


class Module:
    """The slice of torch7's nn.Module protocol the warp path uses."""

    def __init__(self):
        self.output = torch.empty(0)
        self.gradInput = torch.empty(0)

    def forward(self, input):
        return self.updateOutput(input)

    def backward(self, input, gradOutput, scale=1):
        self.updateGradInput(input, gradOutput)
        return self.gradInput

    def __call__(self, input):
        return self.forward(input)

    def clearState(self):
        self.output = torch.empty(0)
        self.gradInput = torch.empty(0) if not isinstance(self.gradInput, list) else []
        return self


def _add_outer_dim(t: torch.Tensor) -> torch.Tensor:
    return t.view((1,) + tuple(t.shape))  # addOuterDim, lua:44-52


# ------------------------------------------------------------------------------------------------
# nn.BilinearSamplerVolumetric
# ------------------------------------------------------------------------------------------------
class BilinearSamplerVolumetric(Module):
    """``BilinearSamplerVolumetric:updateOutput({inputImages, grids})`` /
    ``:updateGradInput({inputImages, grids}, gradOutput)`` (stn3dtorch/BilinearSamplerVolumetric.lua).

    Extras that do not exist in the Lua class (all default to the reference behaviour):
      gridWidth   4 (the Lua class, lua:34) or 3 (the registered-but-unwrapped BHWD twins, ...c:831-836)
      onlyGrid    skip gradInput[1]  -- the reference's disabled onlyGrid switch (...c:594, ...cu:1079)
      onlyVolume  skip gradInput[2]
      algo / vpt  kernel variant selection for benchmarking (see include/volstn_b200.h)
      layout      "BDHWC" (the Lua class: channels-last volume, ``... x 4`` grid) or "BCDHW": volume
                  ``B x C x D x H x W``, grid ``B x G x D x H x W``, output ``B x C x oD x oH x oW`` -- the tensors the
                  reference feeds through materialised ``nn.Transpose`` copies around the sampler
                  (stn3dtorch/test.lua:17,21; models/alignet.lua:11,14,21).  The kernels read and write the
                  channel-first tensors in place through permuted views; no copy is made (CUDA tensors only)
      reuseForwardInputs  host tensors only, OPT-IN (default False = the reference: updateGradInput reads the tensors
                  it is given, ...c:587-640): updateGradInput re-uses the device copies updateOutput made of the
                  same (pointer, shape) volume and grid instead of uploading them again.  Only valid if the inputs
                  are not modified in place between the two calls -- the library cannot see such a modification
      hostAsync   host tensors only: both calls return once their copies and kernels are enqueued; the results are
                  in ``self.output`` / ``self.gradInput`` after ``host_wait()`` (uploads of the next call then
                  overlap the downloads of this one).  Use pinned tensors, and a module (= a set of output
                  buffers) per call in flight
    """

    def __init__(self, gridWidth: int = 4, onlyGrid: bool = False, onlyVolume: bool = False, algo: int = 0, vpt: int = 0,
                 reuseForwardInputs: bool = False, layout: str = "BDHWC", hostAsync: bool = False):
        super().__init__()
        assert layout in ("BDHWC", "BCDHW")
        self.layout = layout
        self.reuseForwardInputs = reuseForwardInputs
        self.hostAsync = hostAsync
        self.gradInput = []  # lua:23  self.gradInput = {}
        assert gridWidth in (3, 4)
        self.gridWidth = gridWidth
        self.onlyGrid = onlyGrid
        self.onlyVolume = onlyVolume
        self.algo = algo
        self.vpt = vpt

    # lua:26-42
    def check(self, input, gradOutput=None):
        inputImages, grids = input[0], input[1]
        if self.layout == "BDHWC":
            assert inputImages.is_contiguous(), "Input images have to be contiguous"
        assert inputImages.dim() == 5
        assert grids.dim() == 5
        assert inputImages.size(0) == grids.size(0)  # batch
        assert grids.size(4) == self.gridWidth  # coordinates
        if gradOutput is not None:
            assert grids.size(0) == gradOutput.size(0)
            assert grids.size(1) == gradOutput.size(1)
            assert grids.size(2) == gradOutput.size(2)
            assert grids.size(3) == gradOutput.size(3)

    def _flags(self) -> int:
        return (self.algo & 0xF00) | _abi.VPT(self.vpt)

    # the nn.Transpose pairs of the reference as views
    @staticmethod
    def _cl(t):
        return t.permute(0, 2, 3, 4, 1)

    @staticmethod
    def _cf(t):
        return t.permute(0, 4, 1, 2, 3)

    def _updateOutput_cf(self, input):
        vol, grid = input[0], input[1]
        assert vol.is_cuda, "layout='BCDHW' is implemented for CUDA tensors"
        assert vol.dim() == 5 and grid.dim() == 5 and vol.is_contiguous(), "Input images have to be contiguous"
        v, g = self._cl(vol), self._cl(grid)
        self.check([v, g])
        _same_place(vol, grid)
        out = torch.empty((vol.size(0), vol.size(1), grid.size(2), grid.size(3), grid.size(4)), dtype=vol.dtype, device=vol.device)
        _require_device(vol.device.index)
        with torch.cuda.device(vol.device):
            rc = _abi.lib().volstn_b200_sampler_forward(_dtype_code(vol), self.gridWidth, _desc(v), _desc(g), _desc(self._cl(out)),
                                                        self._flags(), _stream(vol))
        _abi.check(rc, "BilinearSamplerVolumetric_updateOutput")
        self.output = out
        return out

    def _updateGradInput_cf(self, input, gradOutput):
        vol, grid = input[0], input[1]
        v, g, go = self._cl(vol), self._cl(grid), self._cl(gradOutput)
        self.check([v, g], go)
        _same_place(vol, grid, gradOutput)
        assert gradOutput.size(1) == vol.size(1)
        gv, gg = torch.empty_like(vol), torch.empty_like(grid, memory_format=torch.contiguous_format)
        flags = self._flags() | _abi.BWD_ZERO_GRADVOL
        if self.onlyGrid:
            flags |= _abi.BWD_ONLY_GRID
            gv.zero_()
        if self.onlyVolume:
            flags |= _abi.BWD_ONLY_VOLUME
            gg.zero_()
        _require_device(vol.device.index)
        with torch.cuda.device(vol.device):
            rc = _abi.lib().volstn_b200_sampler_backward(_dtype_code(vol), self.gridWidth, _desc(v), _desc(g), _desc(self._cl(gv)),
                                                         _desc(self._cl(gg)), _desc(go), flags, _stream(vol))
        _abi.check(rc, "BilinearSamplerVolumetric_updateGradInput")
        self.gradInput = [gv, gg]
        return self.gradInput

    # lua:54-81
    def updateOutput(self, input):
        if self.layout == "BCDHW":
            return self._updateOutput_cf(input)
        _inputImages, _grids = input[0], input[1]
        if _inputImages.dim() == 4:
            inputImages, grids = _add_outer_dim(_inputImages), _add_outer_dim(_grids)
        else:
            inputImages, grids = _inputImages, _grids
        self.check([inputImages, grids])
        _same_place(inputImages, grids)
        shape = (inputImages.size(0), grids.size(1), grids.size(2), grids.size(3), inputImages.size(4))
        out = self.output
        if (not isinstance(out, torch.Tensor) or tuple(out.shape) != shape or out.dtype != inputImages.dtype
                or out.device != inputImages.device or not out.is_contiguous()):
            out = torch.empty(shape, dtype=inputImages.dtype, device=inputImages.device,
                              pin_memory=(inputImages.device.type == "cpu" and inputImages.is_pinned()))
        lib = _abi.lib()
        dt = _dtype_code(inputImages)
        dv, dg, do = _desc(inputImages), _desc(grids), _desc(out)
        if inputImages.is_cuda:
            _require_device(inputImages.device.index)
            with torch.cuda.device(inputImages.device):
                rc = lib.volstn_b200_sampler_forward(dt, self.gridWidth, dv, dg, do, self._flags(), _stream(inputImages))
        else:
            # strided grids are honoured like the reference's CPU code honours them (...c:460-473): no copy
            hf = self._flags() | (_abi.HOST_KEEP_INPUTS if self.reuseForwardInputs else 0) | (_abi.HOST_ASYNC if self.hostAsync else 0)
            rc = lib.volstn_b200_host_sampler_forward(_host_ctx(), dt, self.gridWidth, dv, dg, do, hf)
        _abi.check(rc, "BilinearSamplerVolumetric_updateOutput")
        self.output = out[0] if _inputImages.dim() == 4 else out
        return self.output

    # lua:83-117
    def updateGradInput(self, _input, _gradOutput):
        if self.layout == "BCDHW":
            return self._updateGradInput_cf(_input, _gradOutput)
        _inputImages, _grids = _input[0], _input[1]
        if _inputImages.dim() == 4:
            inputImages, grids, gradOutput = (_add_outer_dim(_inputImages), _add_outer_dim(_grids),
                                              _add_outer_dim(_gradOutput))
        else:
            inputImages, grids, gradOutput = _inputImages, _grids, _gradOutput
        self.check([inputImages, grids], gradOutput)
        _same_place(inputImages, grids, gradOutput)
        assert gradOutput.size(4) == inputImages.size(4)
        # lua:101-104  gradInput[i]:resizeAs(input[i]):zero().  The zero-fill of gradInput[1] is
        # delegated to the library (VOLSTN_BWD_ZERO_GRADVOL, one memset node on the same stream);
        # gradInput[2] needs none because the kernel assigns every element (...c:819-822).
        pin = inputImages.device.type == "cpu" and inputImages.is_pinned()
        gi = self.gradInput if isinstance(self.gradInput, list) else []
        while len(gi) < 2:
            gi.append(None)

        def like(cur, ref):
            if (isinstance(cur, torch.Tensor) and cur.dim() == 5 and tuple(cur.shape) == tuple(ref.shape)
                    and cur.dtype == ref.dtype and cur.device == ref.device and cur.is_contiguous()):
                return cur
            return torch.empty(tuple(ref.shape), dtype=ref.dtype, device=ref.device, pin_memory=pin)

        gradInputImages = like(gi[0], inputImages)
        gradGrids = like(gi[1], grids)
        flags = self._flags() | _abi.BWD_ZERO_GRADVOL
        if self.onlyGrid:
            flags |= _abi.BWD_ONLY_GRID
            gradInputImages.zero_()
        if self.onlyVolume:
            flags |= _abi.BWD_ONLY_VOLUME
            gradGrids.zero_()
        lib = _abi.lib()
        dt = _dtype_code(inputImages)
        if inputImages.is_cuda:
            _require_device(inputImages.device.index)
            with torch.cuda.device(inputImages.device):
                rc = lib.volstn_b200_sampler_backward(dt, self.gridWidth, _desc(inputImages), _desc(grids),
                                                      _desc(gradInputImages), _desc(gradGrids), _desc(gradOutput),
                                                      flags, _stream(inputImages))
        else:
            if self.reuseForwardInputs:
                flags |= _abi.HOST_REUSE_INPUTS
            if self.hostAsync:
                flags |= _abi.HOST_ASYNC
            rc = lib.volstn_b200_host_sampler_backward(_host_ctx(), dt, self.gridWidth, _desc(inputImages), _desc(grids),
                                                       _desc(gradInputImages), _desc(gradGrids), _desc(gradOutput), flags)
        _abi.check(rc, "BilinearSamplerVolumetric_updateGradInput")
        if _gradOutput.dim() == 4:
            self.gradInput = [gradInputImages[0], gradGrids[0]]
        else:
            self.gradInput = [gradInputImages, gradGrids]
        return self.gradInput

    def indices(self, inputImages: torch.Tensor, grids: torch.Tensor):
        """Debug dump (no Lua counterpart): floor indices (z0,y0,x0) int32 [V,3] and corner masks uint8 [V]."""
        assert inputImages.is_cuda and grids.is_cuda and grids.dim() == 5
        n = grids.size(0) * grids.size(1) * grids.size(2) * grids.size(3)
        idx = torch.empty((n, 3), dtype=torch.int32, device=grids.device)
        mask = torch.empty((n,), dtype=torch.uint8, device=grids.device)
        vs = (ctypes.c_int64 * 5)(*[int(s) for s in inputImages.shape])
        _require_device(grids.device.index)
        with torch.cuda.device(grids.device):
            rc = _abi.lib().volstn_b200_sampler_indices(_dtype_code(grids), self.gridWidth, vs, _desc(grids),
                                                        idx.data_ptr(), mask.data_ptr(), _stream(grids))
        _abi.check(rc, "sampler_indices")
        return idx, mask


# ------------------------------------------------------------------------------------------------
# nn.VolumetricGridGenerator
# ------------------------------------------------------------------------------------------------
class VolumetricGridGenerator(Module):
    """``nn.VolumetricGridGenerator(depth, height, width)`` (stn3dtorch/VolumetricGridGenerator.lua).

    Field names ``depth / height / width / baseGrid / batchGrid`` are part of the reference's
    serialised form (lua:9-24).  ``baseGrid`` is produced on demand by the forward kernel with
    T = I; the B-replicated ``batchGrid`` (lua:50-55) is never materialised (``batchGrid`` is None).
    """

    def __init__(self, depth: int, height: int, width: int):
        super().__init__()
        assert depth > 1
        assert height > 1
        assert width > 1
        self.depth, self.height, self.width = depth, height, width
        self.batchGrid = None
        self._baseGrid = None

    @property
    def baseGrid(self) -> torch.Tensor:
        if self._baseGrid is None:
            eye = torch.eye(4, dtype=torch.float32, device="cuda").view(1, 4, 4)
            self._baseGrid = self.updateOutput(eye)[0].clone()
        return self._baseGrid

    # lua:37-65
    def updateOutput(self, _transformMatrix):
        transformMatrix = _transformMatrix.view(1, 4, 4) if _transformMatrix.dim() == 2 else _transformMatrix
        assert (transformMatrix.dim() == 3 and transformMatrix.size(1) == 4 and transformMatrix.size(2) == 4), \
            "please input affine transform matrices (bx4x4)"
        T = transformMatrix if transformMatrix.is_contiguous() else transformMatrix.contiguous()
        B = T.size(0)
        shape = (B, self.depth, self.height, self.width, 4)
        out = torch.empty(shape, dtype=T.dtype, device=T.device, pin_memory=(T.device.type == "cpu" and T.is_pinned()))
        lib = _abi.lib()
        dt = _dtype_code(T)
        if T.is_cuda:
            _require_device(T.device.index)
            with torch.cuda.device(T.device):
                rc = lib.volstn_b200_gridgen_forward(dt, T.data_ptr(), out.data_ptr(), B, self.depth, self.height,
                                                     self.width, _stream(T))
        else:
            rc = lib.volstn_b200_host_gridgen_forward(_host_ctx(), dt, T.data_ptr(), out.data_ptr(), B, self.depth,
                                                      self.height, self.width)
        _abi.check(rc, "VolumetricGridGenerator_updateOutput")
        self.output = out[0] if _transformMatrix.dim() == 2 else out
        return self.output

    # lua:67-90
    def updateGradInput(self, _transformMatrix, _gradGrid):
        if _transformMatrix.dim() == 2:
            transformMatrix, gradGrid = _transformMatrix.view(1, 4, 4), _add_outer_dim(_gradGrid)
        else:
            transformMatrix, gradGrid = _transformMatrix, _gradGrid
        B = transformMatrix.size(0)
        assert tuple(gradGrid.shape) == (B, self.depth, self.height, self.width, 4)
        _same_place(transformMatrix, gradGrid)
        gg = gradGrid if gradGrid.is_contiguous() else gradGrid.contiguous()  # :view() in lua:77 needs it too
        gradT = torch.empty((B, 4, 4), dtype=gg.dtype, device=gg.device)
        lib = _abi.lib()
        dt = _dtype_code(gg)
        if gg.is_cuda:
            _require_device(gg.device.index)
            ws_bytes = lib.volstn_b200_gridgen_backward_workspace(B, self.depth, self.height, self.width)
            ws = torch.empty((max(int(ws_bytes), 8) + 7) // 8, dtype=torch.float64, device=gg.device)
            with torch.cuda.device(gg.device):
                rc = lib.volstn_b200_gridgen_backward(dt, gg.data_ptr(), gradT.data_ptr(), B, self.depth, self.height,
                                                      self.width, ws.data_ptr(), ws.numel() * 8, _stream(gg))
        else:
            rc = lib.volstn_b200_host_gridgen_backward(_host_ctx(), dt, gg.data_ptr(), gradT.data_ptr(), B, self.depth,
                                                       self.height, self.width)
        _abi.check(rc, "VolumetricGridGenerator_updateGradInput")
        self.gradInput = gradT[0] if _transformMatrix.dim() == 2 else gradT
        return self.gradInput


# ------------------------------------------------------------------------------------------------
# nn.Cumsum
# ------------------------------------------------------------------------------------------------
class Cumsum(Module):
    """``nn.Cumsum()`` (Cumsum.lua): channel i of ``B x N x d1 x .. x dN`` is prefix-summed along
    spatial axis i; backward is the suffix sum.

    Reference quirk kept by default: ``yxz[i]:squeeze()`` (Cumsum.lua:18) also drops a batch
    dimension of size 1, so the Lua module fails for B = 1 (it scans the wrong axis for i < N and
    raises for i = N).  ``allowSingletonBatch=True`` computes the documented semantics instead.
    """

    def __init__(self, allowSingletonBatch: bool = False):
        super().__init__()
        self.allowSingletonBatch = allowSingletonBatch
        self._gradOutput = None

    def _call(self, x: torch.Tensor, reverse: bool) -> torch.Tensor:
        if x.dim() < 3 or x.size(1) != x.dim() - 2:
            raise ValueError("nn.Cumsum expects B x N x d1 x ... x dN (Cumsum.lua:3-6)")
        if x.size(0) == 1 and not self.allowSingletonBatch:
            raise ValueError("nn.Cumsum: batch of 1 is broken in the reference (squeeze(), Cumsum.lua:18); "
                             "pass allowSingletonBatch=True for the documented semantics")
        x = x if x.is_contiguous() else x.contiguous()  # Cumsum.lua:26-30
        out = torch.empty_like(x)
        N = x.size(1)
        dims = (ctypes.c_int64 * N)(*[int(s) for s in x.shape[2:]])
        lib = _abi.lib()
        dt = _dtype_code(x)
        if x.is_cuda:
            _require_device(x.device.index)
            fn = lib.volstn_b200_cumsum_backward if reverse else lib.volstn_b200_cumsum_forward
            with torch.cuda.device(x.device):
                rc = fn(dt, x.data_ptr(), out.data_ptr(), x.size(0), N, dims, _stream(x))
        else:
            fn = lib.volstn_b200_host_cumsum_backward if reverse else lib.volstn_b200_host_cumsum_forward
            rc = fn(_host_ctx(), dt, x.data_ptr(), out.data_ptr(), x.size(0), N, dims)
        _abi.check(rc, "Cumsum")
        return out

    def updateOutput(self, input):
        self.output = self._call(input, reverse=False)
        return self.output

    def updateGradInput(self, input, gradOutput):
        assert tuple(gradOutput.shape) == tuple(input.shape)
        self.gradInput = self._call(gradOutput, reverse=True)
        return self.gradInput

    def clearState(self):
        self._gradOutput = None  # Cumsum.lua:53-56
        return super().clearState()


# ------------------------------------------------------------------------------------------------
# fused modules: no Lua counterpart; each replaces a sub-graph of reference
# modules and returns that sub-graph's values
# ------------------------------------------------------------------------------------------------
def _vol_view(t: torch.Tensor, layout: str) -> torch.Tensor:
    return t if layout == "BDHWC" else t.permute(0, 2, 3, 4, 1)


def _reuse_or_new(cur, shape, like: torch.Tensor) -> torch.Tensor:
    """``self.output:resize(...)`` of the Lua modules: keep the module's buffer when it already fits (for host
    tensors this also keeps the pinned allocation)."""
    if (isinstance(cur, torch.Tensor) and tuple(cur.shape) == tuple(shape) and cur.dtype == like.dtype
            and cur.device == like.device and cur.is_contiguous()):
        return cur
    return torch.empty(shape, dtype=like.dtype, device=like.device, pin_memory=(not like.is_cuda and like.is_pinned()))


class AffineSamplerVolumetric(Module):
    """``nn.ParallelTable():add(volume):add(nn.VolumetricGridGenerator(depth, height, width))`` followed by
    ``nn.BilinearSamplerVolumetric()`` -- the *projector* of stn3dtorch/test.lua:15-25 and the loader-side affine
    augmentation of dataset_2d.lua:285-311 -- as ONE kernel: the sampling grid ``T[b] . (zn, yn, xn, 1)`` is
    formed in registers and never written (16 B/voxel saved in each direction, no B-replicated base grid).

    ``updateOutput({volume, T})``, ``updateGradInput({volume, T}, gradOutput) -> {gradVolume, gradT}``.
    Values are bit-identical to the two-module chain of this package (gradT: same deterministic reduction
    design, compared within 1e-5 relative)."""

    def __init__(self, depth: int, height: int, width: int, layout: str = "BDHWC", onlyGrid: bool = False, algo: int = 0):
        super().__init__()
        assert depth > 1 and height > 1 and width > 1  # VolumetricGridGenerator.lua:6-8
        assert layout in ("BDHWC", "BCDHW")
        self.depth, self.height, self.width, self.layout, self.onlyGrid = depth, height, width, layout, onlyGrid
        self.algo = algo  # _abi.ALGO_ROW forces the row kernel in the forward (A/B timing)
        self.gradInput = []

    def _T(self, T):
        T = T.view(1, 4, 4) if T.dim() == 2 else T
        assert T.dim() == 3 and T.size(1) == 4 and T.size(2) == 4, "please input affine transform matrices (bx4x4)"
        return T if T.is_contiguous() else T.contiguous()

    def updateOutput(self, input):
        vol, T = input[0], self._T(input[1])
        assert vol.dim() == 5 and vol.is_contiguous() and vol.size(0) == T.size(0)
        _same_place(vol, T)
        v = _vol_view(vol, self.layout)
        B, C = v.size(0), v.size(4)
        shape = (B, self.depth, self.height, self.width, C) if self.layout == "BDHWC" else (B, C, self.depth, self.height, self.width)
        out = _reuse_or_new(self.output, shape, vol)
        if vol.is_cuda:
            _require_device(vol.device.index)
            with torch.cuda.device(vol.device):
                rc = _abi.lib().volstn_b200_affine_sampler_forward(_dtype_code(vol), T.data_ptr(), _desc(v),
                                                                   _desc(_vol_view(out, self.layout)), self.algo & 0xF00,
                                                                   _stream(vol))
        else:  # host tensors (the loader-side augmentation of dataset_2d.lua:285-311): staged through the GPU
            assert self.layout == "BDHWC", "host tensors: channels-last only"
            rc = _abi.lib().volstn_b200_host_affine_sampler_forward(_host_ctx(), _dtype_code(vol), T.data_ptr(), _desc(v),
                                                                    _desc(out), 0)
        _abi.check(rc, "AffineSamplerVolumetric_updateOutput")
        self.output = out
        return out

    def updateGradInput(self, input, gradOutput):
        vol, T = input[0], self._T(input[1])
        _same_place(vol, T, gradOutput)
        v, go = _vol_view(vol, self.layout), _vol_view(gradOutput, self.layout)
        B = v.size(0)
        assert tuple(go.shape) == (B, self.depth, self.height, self.width, v.size(4))
        gv = torch.empty_like(vol)
        gradT = torch.empty((B, 4, 4), dtype=vol.dtype, device=vol.device)
        lib = _abi.lib()
        flags = _abi.BWD_ZERO_GRADVOL
        if self.onlyGrid:
            flags |= _abi.BWD_ONLY_GRID
            gv.zero_()
        if vol.is_cuda:
            ws_bytes = int(lib.volstn_b200_affine_sampler_backward_workspace(B, self.depth, self.height, self.width))
            ws = torch.empty((max(ws_bytes, 8) + 7) // 8, dtype=torch.float64, device=vol.device)
            _require_device(vol.device.index)
            with torch.cuda.device(vol.device):
                rc = lib.volstn_b200_affine_sampler_backward(_dtype_code(vol), T.data_ptr(), _desc(v),
                                                             _desc(_vol_view(gv, self.layout)), gradT.data_ptr(), _desc(go),
                                                             ws.data_ptr(), ws.numel() * 8, flags, _stream(vol))
        else:
            assert self.layout == "BDHWC", "host tensors: channels-last only"
            goc = go if go.is_contiguous() else go.contiguous()
            rc = lib.volstn_b200_host_affine_sampler_backward(_host_ctx(), _dtype_code(vol), T.data_ptr(), _desc(v), _desc(gv),
                                                              gradT.data_ptr(), _desc(goc), flags)
        _abi.check(rc, "AffineSamplerVolumetric_updateGradInput")
        self.gradInput = [gv, gradT[0] if input[1].dim() == 2 else gradT]
        return self.gradInput


def _zeros_like_view(t: torch.Tensor) -> torch.Tensor:
    """an all-zero tensor of t's shape that occupies one element (stride 0 everywhere): read-only by construction --
    torch refuses in-place writes to it"""
    return torch.zeros((), dtype=t.dtype, device=t.device).expand(t.shape)


class CageWarpVolumetric(Module):
    """ALIGNet's warp sub-graph (models/alignet_2d.lua:136-147 with alignet.lua:4-39, lifted to 3-D) as ONE kernel:

        cage  B x 3 x cD x cH x cW  (the output of nn.Cumsum + beta, models/alignet_2d.lua:104-115)
          -> bilinearSampler()({cage, fieldI})      up-sample to the full-resolution field
          -> bilinearSampler()({source, field})     warp the source

    The cage lives in shared memory; the full-resolution field, the constant identity field ``fieldI`` and, in the
    backward pass, the field's gradient never exist in HBM (128 of the chain's 148 B/voxel).

    ``updateOutput({source, cage})``, ``updateGradInput({source, cage}, gradOutput) -> {gradSource, gradCage}``.
    ``onlyGrid=True`` skips gradSource (the source is data behind nn.Identity, models/alignet_2d.lua:23-24)."""

    def __init__(self, depth: Optional[int] = None, height: Optional[int] = None, width: Optional[int] = None,
                 layout: str = "BDHWC", onlyGrid: bool = False, fastArith: bool = False, algo: int = 0):
        super().__init__()
        assert layout in ("BDHWC", "BCDHW")
        self.depth, self.height, self.width, self.layout, self.onlyGrid = depth, height, width, layout, onlyGrid
        self.algo = algo  # _abi.ALGO_ROW keeps the row kernels where the brick kernels would run (A/B timing and tests)
        # VOLSTN_FAST_ARITH (single-channel volumes): separable cage up-sample, FMA sums, factored gradGrid -- values within
        # 1e-5 of the reference chain instead of bit-identical (include/volstn_b200.h)
        self.fastArith = fastArith
        self.gradInput = []

    def _fast(self) -> int:
        return (_abi.FAST_ARITH if self.fastArith else 0) | (self.algo & 0xF00)

    def _out_dims(self, v):
        return (self.depth or v.size(1), self.height or v.size(2), self.width or v.size(3))

    @staticmethod
    def _cage(cage):
        assert cage.dim() == 5 and cage.size(1) == 3, "cage must be B x 3 x cD x cH x cW"
        return cage if cage.is_contiguous() else cage.contiguous()

    def updateOutput(self, input):
        vol, cage = input[0], self._cage(input[1])
        assert vol.dim() == 5 and vol.is_contiguous() and vol.size(0) == cage.size(0)
        _same_place(vol, cage)
        v = _vol_view(vol, self.layout)
        B, C = v.size(0), v.size(4)
        d, h, w = self._out_dims(v)
        out = _reuse_or_new(self.output, (B, d, h, w, C) if self.layout == "BDHWC" else (B, C, d, h, w), vol)
        dims = (ctypes.c_int64 * 3)(*[int(x) for x in cage.shape[2:]])
        if vol.is_cuda:
            _require_device(vol.device.index)
            with torch.cuda.device(vol.device):
                rc = _abi.lib().volstn_b200_cage_warp_forward(_dtype_code(vol), cage.data_ptr(), dims, _desc(v),
                                                              _desc(_vol_view(out, self.layout)), self._fast(), _stream(vol))
        else:
            assert self.layout == "BDHWC", "host tensors: channels-last only"
            rc = _abi.lib().volstn_b200_host_cage_warp_forward(_host_ctx(), _dtype_code(vol), cage.data_ptr(), dims, _desc(v),
                                                               _desc(out), self._fast())
        _abi.check(rc, "CageWarpVolumetric_updateOutput")
        self.output = out
        return out

    def updateGradInput(self, input, gradOutput):
        vol, cage = input[0], self._cage(input[1])
        _same_place(vol, cage, gradOutput)
        v, go = _vol_view(vol, self.layout), _vol_view(gradOutput, self.layout)
        d, h, w = self._out_dims(v)
        assert tuple(go.shape) == (v.size(0), d, h, w, v.size(4))
        # onlyGrid: gradSource is all zeros by definition -- a broadcast zero (no 4-bytes-per-voxel fill per call; the
        # kernels never touch it).  Host tensors keep a real buffer (the host entry points copy into it).
        lazy_zero = self.onlyGrid and vol.is_cuda
        gv = _zeros_like_view(vol) if lazy_zero else torch.empty_like(vol)
        gcage = torch.empty_like(cage)
        dims = (ctypes.c_int64 * 3)(*[int(x) for x in cage.shape[2:]])
        flags = _abi.BWD_ZERO_GRADVOL | self._fast()
        if self.onlyGrid:
            flags |= _abi.BWD_ONLY_GRID
            if not lazy_zero:
                gv.zero_()
        if vol.is_cuda:
            _require_device(vol.device.index)
            with torch.cuda.device(vol.device):
                rc = _abi.lib().volstn_b200_cage_warp_backward(_dtype_code(vol), cage.data_ptr(), dims, _desc(v),
                                                               None if lazy_zero else _desc(_vol_view(gv, self.layout)),
                                                               gcage.data_ptr(), _desc(go), flags, _stream(vol))
        else:
            assert self.layout == "BDHWC", "host tensors: channels-last only"
            goc = go if go.is_contiguous() else go.contiguous()
            rc = _abi.lib().volstn_b200_host_cage_warp_backward(_host_ctx(), _dtype_code(vol), cage.data_ptr(), dims, _desc(v),
                                                                _desc(gv), gcage.data_ptr(), _desc(goc), flags)
        _abi.check(rc, "CageWarpVolumetric_updateGradInput")
        self.gradInput = [gv, gcage]
        return self.gradInput


class CageWarpSmoothL1Volumetric(Module):
    """One training step of ALIGNet's warp head in ONE kernel:

        output   = CageWarpVolumetric:forward({source, cage})
        loss     = nn.SmoothL1Criterion():forward(output, target)          (model.lua:34, train.lua:123)
        gradOut  = criterion:backward(output, target)
        gradInput = CageWarpVolumetric:backward({source, cage}, gradOut)   (train.lua:125-126)
        iou      = computeIOU(output, target)                              (util.lua:110-120)

    Neither ``output`` nor ``gradOut`` has to exist: the kernel reads the source gathers and the target
    (8 B/voxel).  ``forwardBackward({source, cage}, target)`` returns ``(loss, gradInput)``; ``loss`` is a 0-dim
    float64 CUDA tensor (no host synchronisation), ``self.lossPerSample`` the per-sample sums, ``self.iouCounts``
    the exact ``B x 2`` (intersection, union) counts, ``self.output`` the warped volume if ``keepOutput``.
    Single-channel ``B x D x H x W x 1`` CUDA float volumes."""

    def __init__(self, sizeAverage: bool = True, onlyGrid: bool = True, keepOutput: bool = False, fastArith: bool = False,
                 algo: int = 0):
        super().__init__()
        self.sizeAverage, self.onlyGrid, self.keepOutput, self.fastArith = sizeAverage, onlyGrid, keepOutput, fastArith
        self.algo = algo  # _abi.ALGO_ROW: the row kernels (A/B timing and tests)
        self.gradInput = []
        self.loss = self.lossPerSample = self.iouCounts = None

    def forwardBackward(self, input, target):
        vol, cage = input[0], CageWarpVolumetric._cage(input[1])
        assert vol.is_cuda and vol.dim() == 5 and vol.is_contiguous() and vol.size(4) == 1 and vol.size(0) == cage.size(0)
        assert target.dim() == 5 and target.size(0) == vol.size(0) and target.size(4) == 1
        _same_place(vol, cage, target)
        B = vol.size(0)
        n = target.numel()
        out = torch.empty(tuple(target.shape), dtype=vol.dtype, device=vol.device).as_strided(target.shape, target.stride()) \
            if self.keepOutput and target.is_contiguous() else None
        if self.keepOutput and out is None:
            target = target.contiguous()
            out = torch.empty_like(target)
        gv = _zeros_like_view(vol) if self.onlyGrid else torch.empty_like(vol)   # onlyGrid: a broadcast zero, never written
        gcage = torch.empty_like(cage)
        loss_sum = torch.empty((B,), dtype=torch.float64, device=vol.device)
        iou = torch.empty((B, 2), dtype=torch.int64, device=vol.device)
        # THNN: norm = sizeAverage ? 1. / ((real)nElement) : 1.  (double division, stored in `real`)
        norm = ctypes.c_float(1.0 / ctypes.c_float(float(n)).value).value if self.sizeAverage else 1.0
        dims = (ctypes.c_int64 * 3)(*[int(x) for x in cage.shape[2:]])
        flags = _abi.BWD_ZERO_GRADVOL | (_abi.FAST_ARITH if self.fastArith else 0) | (self.algo & 0xF00)
        if self.onlyGrid:
            flags |= _abi.BWD_ONLY_GRID
        _require_device(vol.device.index)
        with torch.cuda.device(vol.device):
            rc = _abi.lib().volstn_b200_cage_warp_step(_dtype_code(vol), cage.data_ptr(), dims, _desc(vol), _desc(target),
                                                       _desc(out) if out is not None else None,
                                                       None if self.onlyGrid else _desc(gv), gcage.data_ptr(),
                                                       loss_sum.data_ptr(), iou.data_ptr(), ctypes.c_float(norm), flags,
                                                       _stream(vol))
        _abi.check(rc, "CageWarpSmoothL1Volumetric_forwardBackward")
        self.lossPerSample, self.iouCounts = loss_sum, iou
        self.loss = loss_sum.sum() / n if self.sizeAverage else loss_sum.sum()
        self.output = out if out is not None else torch.empty(0)
        self.gradInput = [gv, gcage]
        return self.loss, self.gradInput

    def iou(self):
        """Mean over the batch of PER-SAMPLE intersection / union from the kernel's exact counts.  NOT the number
        util.lua's computeIOU prints: that function groups elements by flat index modulo B (see ``computeIOU`` below,
        which reproduces it on a kept output)."""
        c = self.iouCounts.to(torch.float64)
        return (c[:, 0] / c[:, 1]).sum() / c.size(0)


def computeIOU(predTarget: torch.Tensor, gtTarget: torch.Tensor) -> torch.Tensor:
    """``computeIOU(predTarget, gtTarget)`` of util.lua:110-120, operation for operation (plain tensor ops, as in the
    reference; not a kernel of this library).  Inputs are channel-first ``B x C x ...`` like the reference's; only the
    first channel is used (lua:113-114).  Note the grouping the Lua code really performs: ``:view(-1, B)`` of a B-major
    tensor followed by a sum over dim 1 collects, in column c, the elements whose flat index is congruent to c modulo B
    -- not sample c.  The fused step kernels count per sample instead (``iouCounts`` / ``iou()``), a deliberate
    deviation documented in DESIGN.md."""
    B = predTarget.size(0)
    t = (gtTarget[:, :1] > 0.5).double()
    et = (predTarget[:, :1] > 0.5).double()
    t_et_sum = t + et
    intrsct_v = (t_et_sum == 2).double().contiguous().view(-1, B).sum(0, keepdim=True)
    un_v = (t_et_sum > 0).double().contiguous().view(-1, B).sum(0, keepdim=True)
    return (intrsct_v / un_v).sum() / B


class BilinearSamplerSmoothL1Volumetric(Module):
    """``nn.BilinearSamplerVolumetric`` + ``nn.SmoothL1Criterion`` + ``computeIOU`` as one training-step kernel, for
    models that emit a dense sampling field: ``forwardBackward({volume, grids}, target) -> (loss, {gradVolume, gradGrids})``.
    Same conventions as :class:`CageWarpSmoothL1Volumetric`; single-channel CUDA float volumes, ``B x D x H x W x G`` grids."""

    def __init__(self, gridWidth: int = 4, sizeAverage: bool = True, onlyGrid: bool = True, keepOutput: bool = False):
        super().__init__()
        assert gridWidth in (3, 4)
        self.gridWidth, self.sizeAverage, self.onlyGrid, self.keepOutput = gridWidth, sizeAverage, onlyGrid, keepOutput
        self.gradInput = []
        self.loss = self.lossPerSample = self.iouCounts = None

    def forwardBackward(self, input, target):
        vol, grid = input[0], input[1]
        assert vol.is_cuda and vol.dim() == 5 and vol.is_contiguous() and vol.size(4) == 1
        assert grid.dim() == 5 and grid.size(0) == vol.size(0) and grid.size(4) == self.gridWidth
        assert target.dim() == 5 and tuple(target.shape[:4]) == tuple(grid.shape[:4]) and target.size(4) == 1
        _same_place(vol, grid, target)
        B, n = vol.size(0), target.numel()
        target = target if target.is_contiguous() else target.contiguous()
        out = torch.empty_like(target) if self.keepOutput else None
        gv = torch.empty_like(vol)
        gg = torch.empty(tuple(grid.shape), dtype=grid.dtype, device=grid.device)
        loss_sum = torch.empty((B,), dtype=torch.float64, device=vol.device)
        iou = torch.empty((B, 2), dtype=torch.int64, device=vol.device)
        norm = float(torch.tensor(1.0 / float(torch.tensor(float(n), dtype=torch.float32)), dtype=torch.float32)) if self.sizeAverage else 1.0
        flags = _abi.BWD_ZERO_GRADVOL
        if self.onlyGrid:
            flags |= _abi.BWD_ONLY_GRID
            gv.zero_()
        _require_device(vol.device.index)
        with torch.cuda.device(vol.device):
            rc = _abi.lib().volstn_b200_sampler_step(_dtype_code(vol), self.gridWidth, _desc(vol), _desc(grid), _desc(target),
                                                     _desc(out) if out is not None else None, _desc(gv), _desc(gg),
                                                     loss_sum.data_ptr(), iou.data_ptr(), ctypes.c_float(norm), flags,
                                                     _stream(vol))
        _abi.check(rc, "BilinearSamplerSmoothL1Volumetric_forwardBackward")
        self.lossPerSample, self.iouCounts = loss_sum, iou
        self.loss = loss_sum.sum() / n if self.sizeAverage else loss_sum.sum()
        self.output = out if out is not None else torch.empty(0)
        self.gradInput = [gv, gg]
        return self.loss, self.gradInput

    iou = CageWarpSmoothL1Volumetric.iou
