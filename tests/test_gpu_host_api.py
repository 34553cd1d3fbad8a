"""HOST-buffer entry points (the torch.FloatTensor path of `libvolstn`, init.c:13-21): host tensors
in, host tensors out, computed on the GPU in pipelined batch slices.  Must equal the oracle exactly
like the device path does."""
import numpy as np
import pytest
import torch

from alignet_b200 import _abi, nn as vnn, synth
from oracle import oracle
from tests._util import bits_equal, max_rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("B,N,C", [(1, 9, 2), (7, 16, 1), (40, 32, 4)])
def test_host_sampler_matches_oracle(pinned, B, N, C):
    rng = np.random.default_rng(B * 100 + N + C)
    vol = synth.volume("v1", B, N, N, N, C, rng)
    grid = synth.grid("g4", B, N, N, N, rng)
    go = synth.grad_output(B, N, N, N, C, rng)
    tv, tg, tgo = torch.from_numpy(vol), torch.from_numpy(grid), torch.from_numpy(go)
    if pinned:
        tv, tg, tgo = tv.pin_memory(), tg.pin_memory(), tgo.pin_memory()
    m = vnn.BilinearSamplerVolumetric()
    out = m.updateOutput([tv, tg])
    assert not out.is_cuda
    ref = oracle.sampler_forward(vol, grid)
    assert bits_equal(out.numpy(), ref)
    gv, gg = m.updateGradInput([tv, tg], tgo)
    rgv, rgg = oracle.sampler_backward(vol, grid, go)
    assert bits_equal(gg.numpy(), rgg)
    assert max_rel(gv.numpy(), rgv) <= 1e-4


def test_host_double_and_g3():
    rng = np.random.default_rng(5)
    vol = synth.volume("v1", 3, 6, 7, 8, 2, rng, np.float64)
    grid = synth.grid("g3", 3, 5, 6, 7, rng, np.float64, G=3)
    go = synth.grad_output(3, 5, 6, 7, 2, rng, np.float64)
    m = vnn.BilinearSamplerVolumetric(gridWidth=3)
    out = m.updateOutput([torch.from_numpy(vol), torch.from_numpy(grid)])
    assert bits_equal(out.numpy(), oracle.sampler_forward(vol, grid, 3))
    gv, gg = m.updateGradInput([torch.from_numpy(vol), torch.from_numpy(grid)], torch.from_numpy(go))
    rgv, rgg = oracle.sampler_backward(vol, grid, go, 3)
    assert bits_equal(gg.numpy(), rgg) and max_rel(gv.numpy(), rgv) <= 1e-12


def test_host_gridgen_and_cumsum():
    rng = np.random.default_rng(6)
    T = rng.standard_normal((5, 4, 4)).astype(np.float32)
    m = vnn.VolumetricGridGenerator(16, 16, 16)
    out = m.updateOutput(torch.from_numpy(T))
    assert not out.is_cuda
    assert np.max(np.abs(out.numpy() - oracle.gridgen_forward(T, 16, 16, 16))) <= 1e-5
    gg = rng.standard_normal((5, 16, 16, 16, 4)).astype(np.float32)
    gt = m.updateGradInput(torch.from_numpy(T), torch.from_numpy(gg))
    assert max_rel(gt.numpy(), oracle.gridgen_backward(gg)) <= 1e-5
    x = rng.standard_normal((6, 3, 8, 8, 8)).astype(np.float32)
    c = vnn.Cumsum()
    assert bits_equal(c.updateOutput(torch.from_numpy(x)).numpy(), oracle.cumsum_forward(x))
    assert bits_equal(c.updateGradInput(torch.from_numpy(x), torch.from_numpy(x)).numpy(), oracle.cumsum_backward(x))


def test_host_default_rereads_inputs_like_the_reference():
    """The reference's updateGradInput reads the tensors it is given (...c:587-640).  DEFAULT modules must therefore see
    an in-place edit made between updateOutput and updateGradInput; the opt-in reuse flag is the caller's promise
    that there is none."""
    rng = np.random.default_rng(9)
    B, N, C = 9, 12, 2
    vol = torch.from_numpy(synth.volume("v1", B, N, N, N, C, rng)).pin_memory()
    grid = torch.from_numpy(synth.grid("g4", B, N, N, N, rng)).pin_memory()
    go = torch.from_numpy(synth.grad_output(B, N, N, N, C, rng)).pin_memory()
    m = vnn.BilinearSamplerVolumetric()
    assert m.reuseForwardInputs is False
    m.updateOutput([vol, grid])
    grid[..., :3] *= 0.5                                       # in-place edit after forward, same pointer
    vol[0] += 1.0
    rgv, rgg = oracle.sampler_backward(vol.numpy(), grid.numpy(), go.numpy())
    gv, gg = m.updateGradInput([vol, grid], go)
    assert bits_equal(gg.numpy(), rgg) and max_rel(gv.numpy(), rgv) <= 1e-4


def test_host_backward_reuses_forward_inputs_when_asked():
    """VOLSTN_HOST_KEEP_INPUTS / _REUSE_INPUTS (opt-in): the backward call skips the upload of volume and grid when it
    follows a forward call on the same host tensors; any other tensor falls back to uploading."""
    rng = np.random.default_rng(9)
    B, N, C = 9, 12, 2
    vol = torch.from_numpy(synth.volume("v1", B, N, N, N, C, rng)).pin_memory()
    grid = torch.from_numpy(synth.grid("g4", B, N, N, N, rng)).pin_memory()
    go = torch.from_numpy(synth.grad_output(B, N, N, N, C, rng)).pin_memory()
    rgv, rgg = oracle.sampler_backward(vol.numpy(), grid.numpy(), go.numpy())
    m = vnn.BilinearSamplerVolumetric(reuseForwardInputs=True)
    m.updateOutput([vol, grid])
    gv, gg = m.updateGradInput([vol, grid], go)
    assert bits_equal(gg.numpy(), rgg) and max_rel(gv.numpy(), rgv) <= 1e-4
    # a second backward without a forward in between must not reuse anything stale
    grid2 = grid.clone().pin_memory()
    grid2[..., :3] *= 0.5
    rgv2, rgg2 = oracle.sampler_backward(vol.numpy(), grid2.numpy(), go.numpy())
    gv2, gg2 = m.updateGradInput([vol, grid2], go)           # different tensor -> uploaded
    assert bits_equal(gg2.numpy(), rgg2) and max_rel(gv2.numpy(), rgv2) <= 1e-4
    # forward on (vol, grid), backward on (vol, grid2): key mismatch -> falls back to uploading
    m.updateOutput([vol, grid])
    gv3, gg3 = m.updateGradInput([vol, grid2], go)
    assert bits_equal(gg3.numpy(), rgg2)
    # the hazard the flag documents: same pointer, edited in place -> the STALE forward copy is used
    m.updateOutput([vol, grid])
    grid[..., :3] *= 0.5
    gv4, gg4 = m.updateGradInput([vol, grid], go)
    assert bits_equal(gg4.numpy(), rgg)                        # = gradient at the un-edited grid


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_host_strided_tensors_like_the_cpu_reference(dt):
    """The reference's CPU code honours the strides of dims 0-3 of every tensor (...c:460-473, 607-630).  Host grids,
    gradOutputs and OUTPUTS that are narrowed / strided views go through pitched copies -- no .contiguous() -- and
    only the view's own elements of an output buffer are written."""
    rng = np.random.default_rng(21)
    B, C = 5, 2
    vol = synth.volume("v1", B, 6, 7, 8, C, rng, dt)
    big = synth.grid("g4", B, 7, 9, 22, rng, dt)
    gobig = synth.grad_output(B, 7, 9, 22, C, rng, dt)
    views = {
        "narrow_w": (slice(None), slice(None), slice(None), slice(3, 14)),
        "step_w": (slice(None), slice(None), slice(None), slice(0, 22, 2)),
        "narrow_h_step_d": (slice(None), slice(0, 7, 2), slice(2, 8), slice(None)),
        "batch_step": (slice(0, 5, 2), slice(None), slice(None), slice(1, 21, 3)),
    }
    for name, ix in views.items():
        grid, go = big[ix], gobig[ix]
        v = vol[ix[0]]
        ref = oracle.sampler_forward(np.ascontiguousarray(v), np.ascontiguousarray(grid))
        rgv, rgg = oracle.sampler_backward(np.ascontiguousarray(v), np.ascontiguousarray(grid), np.ascontiguousarray(go))
        tv, tg, tgo = torch.from_numpy(np.ascontiguousarray(v)), torch.from_numpy(big)[ix], torch.from_numpy(gobig)[ix]
        assert not tg.is_contiguous() and not tgo.is_contiguous(), name
        m = vnn.BilinearSamplerVolumetric()
        out = m.updateOutput([tv, tg])
        assert bits_equal(out.numpy(), ref), name
        gv, gg = m.updateGradInput([tv, tg], tgo)
        assert bits_equal(gg.numpy(), rgg), name
        assert max_rel(gv.numpy(), rgv) <= (1e-4 if dt == np.float32 else 1e-12), name
        # strided OUTPUTS through the raw ABI: out / gradGrid are views into larger buffers filled with a sentinel
        from alignet_b200 import _abi
        from alignet_b200.nn import _desc, _dtype_code, _host_ctx
        obuf = torch.full(tuple(gobig.shape), -7.0, dtype=tv.dtype)
        ggbuf = torch.full(tuple(big.shape), -7.0, dtype=tv.dtype)
        oview, ggview = obuf[ix], ggbuf[ix]
        lib = _abi.lib()
        _abi.check(lib.volstn_b200_host_sampler_forward(_host_ctx(), _dtype_code(tv), 4, _desc(tv), _desc(tg), _desc(oview), 0), name)
        assert bits_equal(oview.numpy().copy(), ref), name
        gvb = torch.zeros_like(tv)
        _abi.check(lib.volstn_b200_host_sampler_backward(_host_ctx(), _dtype_code(tv), 4, _desc(tv), _desc(tg), _desc(gvb),
                                                         _desc(ggview), _desc(tgo), 0), name)
        assert bits_equal(ggview.numpy().copy(), rgg), name
        mask = torch.ones(tuple(gobig.shape), dtype=torch.bool)
        mask[ix] = False
        assert bool((obuf[mask] == -7.0).all()), name          # nothing outside the view was written
        gmask = torch.ones(tuple(big.shape), dtype=torch.bool)
        gmask[ix] = False
        assert bool((ggbuf[gmask] == -7.0).all()), name


def test_host_broadcast_grid():
    """An expanded (stride-0 batch) grid -- one field shared by the whole batch -- is read in place."""
    rng = np.random.default_rng(22)
    B, N = 6, 10
    vol = synth.volume("v1", B, N, N, N, 1, rng)
    g1 = synth.grid("g3", 1, N, N, N, rng)
    tg = torch.from_numpy(g1).expand(B, N, N, N, 4)
    out = vnn.BilinearSamplerVolumetric().updateOutput([torch.from_numpy(vol), tg])
    assert bits_equal(out.numpy(), oracle.sampler_forward(vol, np.ascontiguousarray(np.broadcast_to(g1, (B, N, N, N, 4)))))


def test_host_accumulates_into_gradvol_with_flags_zero():
    """flags = 0 is the reference's own contract (...c:730-789: gradVol +=): what an unmodified Lua wrapper gets."""
    from alignet_b200 import _abi
    from alignet_b200.nn import _desc, _host_ctx
    rng = np.random.default_rng(23)
    B, N, C = 4, 9, 3
    vol = synth.volume("v1", B, N, N, N, C, rng)
    grid = synth.grid("g2", B, N, N, N, rng, csz=4)
    go = synth.grad_output(B, N, N, N, C, rng)
    init = rng.standard_normal(vol.shape).astype(np.float32)
    rgv, rgg = oracle.sampler_backward(vol, grid, go, grad_vol=init.copy())
    gv, gg = torch.from_numpy(init.copy()), torch.empty(grid.shape)
    _abi.check(_abi.lib().volstn_b200_host_sampler_backward(_host_ctx(), 0, 4, _desc(torch.from_numpy(vol)), _desc(torch.from_numpy(grid)),
                                                            _desc(gv), _desc(gg), _desc(torch.from_numpy(go)), 0), "bwd")
    assert bits_equal(gg.numpy(), rgg) and max_rel(gv.numpy(), rgv) <= 1e-4


@pytest.mark.parametrize("reuse", [False, True])
def test_host_async_pipeline(reuse):
    """VOLSTN_HOST_ASYNC: calls return after enqueueing; several steps in flight (a module = a set of output buffers per
    step), each completed by its ticket; results equal the oracle's."""
    rng = np.random.default_rng(31)
    B, N = 12, 16
    steps = []
    for i in range(4):
        vol = torch.from_numpy(synth.volume("v1", B, N, N, N, 1, rng)).pin_memory()
        grid = torch.from_numpy(synth.grid("g4" if i % 2 else "g2", B, N, N, N, rng, csz=4)).pin_memory()
        go = torch.from_numpy(synth.grad_output(B, N, N, N, 1, rng)).pin_memory()
        steps.append((vol, grid, go, vnn.BilinearSamplerVolumetric(hostAsync=True, reuseForwardInputs=reuse)))
    tickets = []
    for vol, grid, go, m in steps:
        m.updateOutput([vol, grid])
        m.updateGradInput([vol, grid], go)
        tickets.append(vnn.host_ticket())
    assert tickets == sorted(tickets) and len(set(tickets)) == len(tickets)
    for (vol, grid, go, m), t in zip(steps, tickets):
        vnn.host_wait(t)
        assert bits_equal(m.output.numpy(), oracle.sampler_forward(vol.numpy(), grid.numpy()))
        rgv, rgg = oracle.sampler_backward(vol.numpy(), grid.numpy(), go.numpy())
        assert bits_equal(m.gradInput[1].numpy(), rgg)
        assert max_rel(m.gradInput[0].numpy(), rgv) <= 1e-4
    vnn.host_wait()


def test_host_gridgen_backward_ragged_last_slice():
    """B * blocks-per-sample(B) is not monotone in B, so a shorter last slice can need MORE workspace than a full one
    (64^3, batch 64: slices of 11 and a last slice of 9).  The context's workspace covers every slice length."""
    rng = np.random.default_rng(41)
    N = 64
    for B in (20, 31):
        gg = rng.standard_normal((B, N, N, N, 4)).astype(np.float32)
        T = np.tile(np.eye(4, dtype=np.float32), (B, 1, 1))
        m = vnn.VolumetricGridGenerator(N, N, N)
        gt = m.updateGradInput(torch.from_numpy(T), torch.from_numpy(gg))
        assert max_rel(gt.numpy(), oracle.gridgen_backward(gg)) <= 1e-5
        # the arena's neighbours are intact: a second, different call on the same context still matches
        x = rng.standard_normal((6, 3, 8, 8, 8)).astype(np.float32)
        assert bits_equal(vnn.Cumsum().updateOutput(torch.from_numpy(x)).numpy(), oracle.cumsum_forward(x))


def test_host_calls_from_concurrent_threads():
    """The reference's loader threads call the sampler concurrently on CPU tensors (data.lua:11-24,
    dataset_2d.lua:293-306).  The library is re-entrant with one staging context per thread (nn.py keeps them in
    thread-local storage); device-pointer calls on separate streams are independent as well."""
    import threading
    rng = np.random.default_rng(9)
    jobs = []
    for i in range(6):
        B, N = 3 + i, 12 + i
        vol = synth.volume("v1", B, N, N, N, 1, rng)
        grid = synth.grid("g4", B, N, N, N, rng)
        go = synth.grad_output(B, N, N, N, 1, rng)
        T = np.tile(np.eye(4, dtype=np.float32), (B, 1, 1))
        T[:, :3, 3] = 0.1 * rng.standard_normal((B, 3)).astype(np.float32)
        jobs.append((vol, grid, go, T))
    results, errors = [None] * len(jobs), []

    def work(i):
        try:
            vol, grid, go, T = jobs[i]
            for _ in range(3):  # repeated calls reuse the thread's context
                if i % 2 == 0:  # host tensors
                    m = vnn.BilinearSamplerVolumetric()
                    out = m.updateOutput([torch.from_numpy(vol), torch.from_numpy(grid)]).clone()
                    gv, gg = m.updateGradInput([torch.from_numpy(vol), torch.from_numpy(grid)], torch.from_numpy(go))
                    aff = vnn.AffineSamplerVolumetric(*vol.shape[1:4]).updateOutput([torch.from_numpy(vol), torch.from_numpy(T)])
                else:           # device tensors on this thread's own stream
                    st = torch.cuda.Stream()
                    with torch.cuda.stream(st):
                        m = vnn.BilinearSamplerVolumetric()
                        tv, tg, tgo = torch.from_numpy(vol).cuda(), torch.from_numpy(grid).cuda(), torch.from_numpy(go).cuda()
                        out = m.updateOutput([tv, tg]).clone()
                        gv, gg = m.updateGradInput([tv, tg], tgo)
                        aff = vnn.AffineSamplerVolumetric(*vol.shape[1:4]).updateOutput([tv, torch.from_numpy(T).cuda()])
                    st.synchronize()
                results[i] = (out.cpu().numpy(), gv.cpu().numpy(), gg.cpu().numpy(), aff.cpu().numpy())
        except Exception as e:  # noqa: BLE001
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for i, (vol, grid, go, T) in enumerate(jobs):
        out, gv, gg, aff = results[i]
        assert bits_equal(out, oracle.sampler_forward(vol, grid)), i
        rgv, rgg = oracle.sampler_backward(vol, grid, go)
        assert bits_equal(gg, rgg), i
        assert max_rel(gv, rgv) <= 1e-4, i
        agrid = vnn.VolumetricGridGenerator(*vol.shape[1:4]).updateOutput(torch.from_numpy(T).cuda()).cpu().numpy()
        assert bits_equal(aff, oracle.sampler_forward(vol, agrid)), i


def test_host_pin_unpin_contexts_and_error_reporting():
    """The housekeeping entry points of the host ABI: pin / unpin of caller-owned buffers (cudaHostRegister: optional, the
    results must not depend on it), explicit contexts (create / destroy, more than one alive), placement queries, and the
    per-thread CUDA error record behind VOLSTN_ERR_CUDA."""
    import ctypes
    lib = _abi.lib()
    rng = np.random.default_rng(9)
    B, N, C = 5, 12, 2
    vol = synth.volume("v1", B, N, N, N, C, rng)
    grid = synth.grid("g2", B, N, N, N, rng, csz=4)
    ref = oracle.sampler_forward(vol, grid)
    ctxs = [ctypes.c_void_p() for _ in range(2)]
    for c in ctxs:
        assert lib.volstn_b200_host_ctx_create(0, ctypes.byref(c)) == 0 and c.value
    out_plain, out_pinned = np.empty_like(ref), np.empty_like(ref)
    d = vnn._desc
    tv, tg = torch.from_numpy(vol), torch.from_numpy(grid)
    assert lib.volstn_b200_host_sampler_forward(ctxs[0], 0, 4, d(tv), d(tg), d(torch.from_numpy(out_plain)), 0) == 0
    for a in (vol, grid, out_pinned):
        assert lib.volstn_b200_host_pin(ctypes.c_void_p(a.ctypes.data), ctypes.c_size_t(a.nbytes)) == 0
    assert lib.volstn_b200_host_sampler_forward(ctxs[1], 0, 4, d(tv), d(tg), d(torch.from_numpy(out_pinned)), 0) == 0
    for a in (vol, grid, out_pinned):
        assert lib.volstn_b200_host_unpin(ctypes.c_void_p(a.ctypes.data)) == 0
    assert bits_equal(out_plain, ref) and bits_equal(out_pinned, ref)
    assert lib.volstn_b200_host_unpin(ctypes.c_void_p(vol.ctypes.data)) == _abi.ERR_CUDA          # not registered any more ...
    assert lib.volstn_b200_last_cuda_error() != 0                                                   # ... and the record says why
    assert len(lib.volstn_b200_last_cuda_error_string()) > 0
    assert lib.volstn_b200_host_ctx_create(10 ** 6, ctypes.byref(ctypes.c_void_p())) != 0           # no such device
    node = lib.volstn_b200_host_numa_node(0)
    assert node >= -1
    assert lib.volstn_b200_host_bind_thread(0) >= 0                                                 # a no-op without topology information
    for c in ctxs:
        assert lib.volstn_b200_host_ctx_destroy(c) == 0
