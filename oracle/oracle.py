"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY (CPU checker and CPU baseline).

Nothing under alignet_b200/ imports this module.  It is used by tests/, by
__graft_entry__.smoke() as the checker, and by bench.py's ``cpu_baseline`` / ``--impl reference``
legs as the thing timed on the host cores.

Three CPU implementations live here:

* ``ref``      -- the reference's own ``stn3dtorch/generic/BilinearSamplerVolumetric.c`` compiled
                  UNMODIFIED behind ``oracle/ref_shim.c`` into ``oracle/_ref/`` (float and double).
                  Only buildable where ``/root/reference`` exists; the built ``.so`` travels.
* ``port``     -- ``oracle/volstn_oracle.c``, our plain-C restatement (float and double).  Pinned
                  bit-for-bit against ``ref`` (tests/test_oracle_vs_ref.py) and against the committed
                  golden vectors (tests/test_oracle_golden.py).
* numpy restatements of ``Cumsum.lua`` and ``stn3dtorch/VolumetricGridGenerator.lua``.  Those two
  modules are Lua over Torch7 ``TH`` (``tensor:cumsum``, ``torch.bmm``), which is not vendored in
  /root/reference and has no pinned version (install.sh:5 clones torch/distro at HEAD), so their
  parity is UNPINNED: defined by the Lua text plus TH semantics as documented below.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libvolstn_oracle.so")
REF_SO = {np.dtype(np.float32): os.path.join(HERE, "_ref", "libvolstn_ref_f32.so"),
          np.dtype(np.float64): os.path.join(HERE, "_ref", "libvolstn_ref_f64.so")}


class _Tensor5(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("size", ctypes.c_long * 5), ("stride", ctypes.c_long * 5)]


def _desc(a: np.ndarray) -> _Tensor5:
    assert a.ndim == 5, a.shape
    assert a.strides[4] == a.itemsize or a.shape[4] == 1, "last dimension must have unit stride"
    t = _Tensor5()
    t.data = a.ctypes.data
    for i in range(5):
        t.size[i] = a.shape[i]
        t.stride[i] = a.strides[i] // a.itemsize
    return t


def build(verbose: bool = False) -> None:
    """Compile the C restatement and, when /root/reference is present, the reference itself."""
    r = subprocess.run(["make", "-C", HERE, "oracle", "ref", "refcuda", "refinit"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


_lock = threading.Lock()
_port = None
_ref = {}


def port_lib():
    global _port
    with _lock:
        if _port is None:
            if not os.path.exists(PORT_SO):
                build()
            _port = ctypes.CDLL(PORT_SO)
    return _port


def have_ref() -> bool:
    return all(os.path.exists(p) for p in REF_SO.values())


def ref_lib(dtype):
    dtype = np.dtype(dtype)
    with _lock:
        if dtype not in _ref:
            if not os.path.exists(REF_SO[dtype]):
                raise FileNotFoundError(
                    f"{REF_SO[dtype]} missing: run `make -C oracle ref` where /root/reference exists")
            lib = ctypes.CDLL(REF_SO[dtype])
            assert lib.ref_sizeof_real() == dtype.itemsize
            _ref[dtype] = lib
    return _ref[dtype]


def default_impl() -> str:
    """``ref`` (the reference's own C, compiled unmodified) whenever oracle/_ref is present -- it travels to the
    GPU box with the snapshot -- else the pinned C restatement."""
    return "ref" if have_ref() else "port"


def _suffix(dtype) -> str:
    return {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}[np.dtype(dtype)]


def _check(vol, grid, G):
    assert vol.ndim == 5 and grid.ndim == 5 and vol.shape[0] == grid.shape[0]
    assert grid.shape[4] == G and G in (3, 4)
    assert vol.dtype == grid.dtype and vol.dtype in (np.float32, np.float64)


# --------------------------------------------------------------------------------------------
# sampler: forward / backward / index dump  (impl = "port" | "ref")
# --------------------------------------------------------------------------------------------
def sampler_forward(vol: np.ndarray, grid: np.ndarray, G: int = 4, impl: str = None, out=None) -> np.ndarray:
    """nn_(BilinearSamplerVolumetric_updateOutput) (G=4, ...c:442) / BilinearSamplerBHWD (G=3, ...c:9).
    impl: "ref" | "port" | None = ``default_impl()``."""
    _check(vol, grid, G)
    impl = impl or default_impl()
    if out is None:
        out = np.empty(grid.shape[:4] + (vol.shape[4],), vol.dtype)
    tv, tg, to = _desc(vol), _desc(grid), _desc(out)
    if impl == "port":
        fn = getattr(port_lib(), "oracle_sampler_forward_" + _suffix(vol.dtype))
        rc = fn(ctypes.c_int(G), ctypes.byref(tv), ctypes.byref(tg), ctypes.byref(to))
        assert rc == 0
    else:
        arr = (_Tensor5 * 3)(tv, tg, to)
        rc = ref_lib(vol.dtype).ref_call(0 if G == 4 else 2, 3, arr)
        assert rc == 1
    return out


def sampler_backward(vol, grid, grad_out, G: int = 4, impl: str = None, grad_vol=None, only_grid=False):
    """nn_(BilinearSamplerVolumetric_updateGradInput) (...c:585) / BHWD twin (...c:169).

    Returns (gradVol, gradGrid).  gradVol is zero-filled first, as the Lua wrapper does
    (BilinearSamplerVolumetric.lua:101-104), unless an accumulator is passed in.
    """
    _check(vol, grid, G)
    impl = impl or default_impl()
    assert grad_out.shape == grid.shape[:4] + (vol.shape[4],)
    if grad_vol is None:
        grad_vol = np.zeros_like(vol)
    grad_grid = np.zeros_like(grid)
    # the reference never enables onlyGrid (...c:594): with impl="ref" its gradVol goes to a scratch tensor
    gv_arg = np.zeros_like(vol) if (only_grid and impl != "port") else grad_vol
    ts = [_desc(vol), _desc(grid), _desc(gv_arg), _desc(grad_grid), _desc(grad_out)]
    if impl == "port":
        fn = getattr(port_lib(), "oracle_sampler_backward_" + _suffix(vol.dtype))
        rc = fn(ctypes.c_int(G), *[ctypes.byref(t) for t in ts], ctypes.c_int(int(only_grid)))
        assert rc == 0
    else:
        arr = (_Tensor5 * 5)(*ts)
        rc = ref_lib(vol.dtype).ref_call(1 if G == 4 else 3, 5, arr)
        assert rc == 1
    return grad_vol, grad_grid


def sampler_indices(vol_shape, grid: np.ndarray):
    """(z0,y0,x0) int32 [V,3] and the 8-bit corner mask uint8 [V] -- port only (a debug dump the
    reference does not have; the port's values are pinned through the bit-exact output parity)."""
    n = int(np.prod(grid.shape[:4]))
    idx = np.empty((n, 3), np.int32)
    mask = np.empty((n,), np.uint8)
    size = (ctypes.c_long * 5)(*[int(s) for s in vol_shape])
    tg = _desc(grid)
    fn = getattr(port_lib(), "oracle_sampler_indices_" + _suffix(grid.dtype))
    rc = fn(size, ctypes.byref(tg), idx.ctypes.data_as(ctypes.c_void_p), mask.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return idx, mask


def sampler_forward_backward_threaded(vol, grid, grad_out, G=4, impl="ref", threads=1, bufs=None):
    """Batch-sliced fwd+bwd over `threads` host threads (ctypes releases the GIL).  Valid because
    samples are independent (batch is the outermost loop, ...c:482,641).  Used only as the timed
    CPU baseline.  One call = what the Lua module does per step: updateOutput into self.output (resized once, i.e.
    pre-allocated: pass `bufs`), then updateGradInput = zero-fill of BOTH gradInputs
    (BilinearSamplerVolumetric.lua:101-104) + the C function."""
    B = vol.shape[0]
    threads = max(1, min(threads, B))
    if bufs is None:
        bufs = {}
    if "out" not in bufs:
        bufs["out"] = np.empty(grid.shape[:4] + (vol.shape[4],), vol.dtype)
        bufs["gv"] = np.empty_like(vol)
        bufs["gg"] = np.empty_like(grid)
    out, gv, gg = bufs["out"], bufs["gv"], bufs["gg"]
    bounds = np.linspace(0, B, threads + 1).astype(int)
    lib_call = _backward_into

    def work(i):
        lo, hi = bounds[i], bounds[i + 1]
        if hi <= lo:
            return
        sampler_forward(vol[lo:hi], grid[lo:hi], G, impl, out=out[lo:hi])
        gv[lo:hi] = 0          # lua:103 gradInput[i]:resizeAs(input[i]):zero()
        gg[lo:hi] = 0
        lib_call(vol[lo:hi], grid[lo:hi], grad_out[lo:hi], gv[lo:hi], gg[lo:hi], G, impl)

    if threads == 1:
        work(0)
    else:
        ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    return out, gv, gg


def _backward_into(vol, grid, grad_out, grad_vol, grad_grid, G, impl):
    """the backward C function on caller-owned gradient tensors (no allocation inside the timed baseline)"""
    ts = [_desc(vol), _desc(grid), _desc(grad_vol), _desc(grad_grid), _desc(grad_out)]
    if impl == "port":
        fn = getattr(port_lib(), "oracle_sampler_backward_" + _suffix(vol.dtype))
        rc = fn(ctypes.c_int(G), *[ctypes.byref(t) for t in ts], ctypes.c_int(0))
        assert rc == 0
    else:
        arr = (_Tensor5 * 5)(*ts)
        rc = ref_lib(vol.dtype).ref_call(1 if G == 4 else 3, 5, arr)
        assert rc == 1


# --------------------------------------------------------------------------------------------
# nn.Cumsum  (Cumsum.lua:12-51)  -- parity UNPINNED (TH arithmetic, see module docstring)
# --------------------------------------------------------------------------------------------
def cumsum_forward(x: np.ndarray, accumulate=np.float64) -> np.ndarray:
    """Cumsum.lua:12-22: channel i (0-based) of a B x N x d1..dN tensor is prefix-summed along
    spatial axis i.  TH's CPU ``cumsum`` accumulates float tensors in ``accreal`` = double and
    rounds each prefix to float ([memory]; THTensorMath.c cumsum); THC accumulates in float.
    ``accumulate`` selects which of the two is restated."""
    assert x.ndim >= 3 and x.shape[1] == x.ndim - 2, "input must be B x N x d1 x ... x dN"
    out = np.empty_like(x)
    for i in range(x.shape[1]):
        out[:, i] = np.cumsum(x[:, i].astype(accumulate), axis=1 + i, dtype=accumulate).astype(x.dtype)
    return out


def cumsum_backward(grad_out: np.ndarray, accumulate=np.float64) -> np.ndarray:
    """Cumsum.lua:39-48: flip, cumsum, flip along the same axis  ==  suffix sum."""
    assert grad_out.ndim >= 3 and grad_out.shape[1] == grad_out.ndim - 2
    gin = np.empty_like(grad_out)
    for i in range(grad_out.shape[1]):
        g = np.flip(grad_out[:, i].astype(accumulate), axis=1 + i)
        gin[:, i] = np.flip(np.cumsum(g, axis=1 + i, dtype=accumulate), axis=1 + i).astype(grad_out.dtype)
    return gin


# --------------------------------------------------------------------------------------------
# nn.VolumetricGridGenerator (stn3dtorch/VolumetricGridGenerator.lua) -- parity UNPINNED
# --------------------------------------------------------------------------------------------
def gridgen_base(D: int, H: int, W: int, dtype=np.float32) -> np.ndarray:
    """VolumetricGridGenerator.lua:12-23: (-1 + (k-1)/(n-1)*2) evaluated in Lua doubles, stored in
    the default tensor type; 4th component 1."""
    def axis(n):
        k = np.arange(n, dtype=np.float64)
        return (-1.0 + k / float(n - 1) * 2.0)
    base = np.empty((D, H, W, 4), np.float64)
    base[..., 0] = axis(D)[:, None, None]
    base[..., 1] = axis(H)[None, :, None]
    base[..., 2] = axis(W)[None, None, :]
    base[..., 3] = 1.0
    return base.astype(dtype)


def gridgen_forward(T: np.ndarray, D: int, H: int, W: int, accumulate=np.float64) -> np.ndarray:
    """VolumetricGridGenerator.lua:58-60: out(B,DHW,4) = batchGrid(B,DHW,4) x T^T(B,4,4).
    BLAS summation order over the K=4 inner dimension is unspecified, hence a tolerance, and
    ``accumulate`` brackets it (float64 = exact-ish, float32 = sequential k=0..3)."""
    assert T.ndim == 3 and T.shape[1:] == (4, 4)
    base = gridgen_base(D, H, W, T.dtype).reshape(-1, 4)
    if accumulate == np.float32 and T.dtype == np.float32:
        out = np.zeros((T.shape[0], base.shape[0], 4), np.float32)
        for k in range(4):
            out = out + base[None, :, k, None] * T[:, None, :, k]
    else:
        out = np.einsum("vk,bik->bvi", base.astype(accumulate), T.astype(accumulate))
    return out.astype(T.dtype).reshape(T.shape[0], D, H, W, 4)


def gridgen_backward(grad_grid: np.ndarray, accumulate=np.float64) -> np.ndarray:
    """VolumetricGridGenerator.lua:79-82: gradT(B,4,4) = gradGrid^T(B,4,DHW) x batchGrid(B,DHW,4)."""
    B, D, H, W, four = grad_grid.shape
    assert four == 4
    base = gridgen_base(D, H, W, grad_grid.dtype).reshape(-1, 4)
    gt = np.einsum("bvi,vj->bij", grad_grid.reshape(B, -1, 4).astype(accumulate), base.astype(accumulate))
    return gt.astype(grad_grid.dtype)


# --------------------------------------------------------------------------------------------
# nn.SmoothL1Criterion (model.lua:34) and computeIOU (util.lua:110-120) -- parity UNPINNED: the criterion lives in
# Torch7's THNN (generic/SmoothL1Criterion.c), which is not vendored; restated from its published algorithm:
#   updateOutput   : z = |input - target| ; sum += z < 1 ? 0.5 z^2 : z - 0.5  (accreal = double) ; sum /= nElement
#   updateGradInput: x = input - target ; grad = x < -1 ? -norm : x > 1 ? norm : norm * x ; norm = 1 / nElement
# --------------------------------------------------------------------------------------------
def smooth_l1_forward(inp: np.ndarray, target: np.ndarray, size_average: bool = True) -> float:
    z = np.abs(inp - target).astype(np.float64)       # the subtraction happens in `real`, the sum in double
    s = float(np.sum(np.where(z < 1.0, 0.5 * z * z, z - 0.5)))
    return s / inp.size if size_average else s


def smooth_l1_backward(inp: np.ndarray, target: np.ndarray, size_average: bool = True) -> np.ndarray:
    norm = inp.dtype.type(1.0 / float(inp.dtype.type(inp.size))) if size_average else inp.dtype.type(1.0)
    x = inp - target
    return np.where(x < -1, -norm, np.where(x > 1, norm, norm * x)).astype(inp.dtype)


def iou_reference(pred: np.ndarray, target: np.ndarray) -> float:
    """computeIOU exactly as util.lua:110-120 writes it, including its grouping: the thresholded tensors (B-major,
    first channel) are viewed as ``(-1, B)`` and summed over dim 1, so "column c" collects the elements whose FLAT index
    is congruent to c modulo B -- NOT sample c.  Returns mean_c(intersection_c / union_c)."""
    B = pred.shape[0]
    t = (target.reshape(B, -1) > 0.5).astype(np.float64).reshape(-1, B)     # :view(-1, B) of the contiguous tensor
    e = (pred.reshape(B, -1) > 0.5).astype(np.float64).reshape(-1, B)
    s = t + e
    inter, union = (s == 2).sum(0).astype(np.float64), (s > 0).sum(0).astype(np.float64)
    return float((inter / union).sum() / B)


def iou_counts(pred: np.ndarray, target: np.ndarray):
    """PER-SAMPLE (#(t and et), #(t or et)) with t = target > 0.5, et = pred > 0.5 (the thresholds of util.lua:113-114).
    This is what the fused step kernels count.  It is a DELIBERATE deviation from computeIOU's grouping (see
    iou_reference): per-sample intersection over union is what the metric's name says; the reference's modulo-B
    grouping is reproduced by alignet_b200.nn.computeIOU on the kept output, not by the kernel."""
    B = pred.shape[0]
    t, e = target.reshape(B, -1) > 0.5, pred.reshape(B, -1) > 0.5
    return np.stack([(t & e).sum(1), (t | e).sum(1)], 1).astype(np.uint64)


# --------------------------------------------------------------------------------------------
# the reference's own CUDA kernels recompiled for sm_100a (oracle/ref_cuda_shim.cu): the GPU incumbent.
# Arguments are torch CUDA tensors; launches on the current stream, like the reference (...cu:706).
# --------------------------------------------------------------------------------------------
REF_CUDA_SO = os.path.join(HERE, "_ref", "libcuvolstn_ref.so")


class _CudaTensor5(ctypes.Structure):
    _fields_ = [("size", ctypes.c_long * 5), ("stride", ctypes.c_long * 5), ("data", ctypes.c_void_p)]


def have_ref_cuda() -> bool:
    return os.path.exists(REF_CUDA_SO)


_ref_cuda = None


def _ref_cuda_lib():
    global _ref_cuda
    if _ref_cuda is None:
        _ref_cuda = ctypes.CDLL(REF_CUDA_SO)
        _ref_cuda.ref_cuda_call.restype = ctypes.c_int
        _ref_cuda.ref_cuda_call.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return _ref_cuda


def _cuda_desc(t):
    d = _CudaTensor5()
    d.data = t.data_ptr()
    for i in range(5):
        d.size[i] = t.size(i)
        d.stride[i] = t.stride(i)
    return d


def ref_cuda_call(which: int, tensors) -> None:
    """which: 0 Volumetric fwd (vol, grid, out) / 1 bwd (vol, grid, gradVol, gradGrid, gradOut); 2 / 3 the G = 3
    twins.  The kernels require packed float tensors (...cu:606) and B * D <= 65535 (blockIdx.z, ...cu:702)."""
    import torch
    arr = (_CudaTensor5 * len(tensors))(*[_cuda_desc(t) for t in tensors])
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = _ref_cuda_lib().ref_cuda_call(which, len(tensors), ctypes.cast(arr, ctypes.c_void_p), st)
    if rc != 0:
        raise RuntimeError("reference CUDA launcher reported an error")
