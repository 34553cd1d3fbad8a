"""HOST-buffer entry points (the torch.FloatTensor path of `libvolstn`, init.c:13-21): host tensors
in, host tensors out, computed on the GPU in pipelined batch slices.  Must equal the oracle exactly
like the device path does."""
import numpy as np
import pytest
import torch

from alignet_b200 import nn as vnn, synth
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


def test_host_backward_reuses_forward_inputs():
    """VOLSTN_HOST_KEEP_INPUTS / _REUSE_INPUTS: the backward call skips the upload of volume and grid when it
    follows a forward call on the same host tensors (nn.Module's contract); any other tensor falls back."""
    rng = np.random.default_rng(9)
    B, N, C = 9, 12, 2
    vol = torch.from_numpy(synth.volume("v1", B, N, N, N, C, rng)).pin_memory()
    grid = torch.from_numpy(synth.grid("g4", B, N, N, N, rng)).pin_memory()
    go = torch.from_numpy(synth.grad_output(B, N, N, N, C, rng)).pin_memory()
    rgv, rgg = oracle.sampler_backward(vol.numpy(), grid.numpy(), go.numpy())
    m = vnn.BilinearSamplerVolumetric()          # reuseForwardInputs=True by default
    m.updateOutput([vol, grid])
    gv, gg = m.updateGradInput([vol, grid], go)
    assert bits_equal(gg.numpy(), rgg) and max_rel(gv.numpy(), rgv) <= 1e-4
    # a second backward without a forward in between must not reuse anything stale: change the grid in place
    grid2 = grid.clone().pin_memory()
    grid2[..., :3] *= 0.5
    rgv2, rgg2 = oracle.sampler_backward(vol.numpy(), grid2.numpy(), go.numpy())
    gv2, gg2 = m.updateGradInput([vol, grid2], go)           # different tensor -> uploaded
    assert bits_equal(gg2.numpy(), rgg2) and max_rel(gv2.numpy(), rgv2) <= 1e-4
    # forward on (vol, grid), backward on (vol, grid2): key mismatch -> falls back to uploading
    m.updateOutput([vol, grid])
    gv3, gg3 = m.updateGradInput([vol, grid2], go)
    assert bits_equal(gg3.numpy(), rgg2)
    # opting out re-uploads even for identical tensors (inputs mutated in place between the calls)
    m2 = vnn.BilinearSamplerVolumetric(reuseForwardInputs=False)
    m2.updateOutput([vol, grid])
    grid[..., :3] *= 0.5                                       # in-place edit after forward
    gv4, gg4 = m2.updateGradInput([vol, grid], go)
    assert bits_equal(gg4.numpy(), rgg2)


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
