"""The N > 1 path on CPU: world_size-2 `gloo` processes shard a batch with alignet_b200.shard, run the
(test-only) oracle on their own slice -- there is no CPU compute path in the product, the oracle
stands in for the kernels here -- and the gathered result must equal the unsharded one bit for bit,
because samples are independent (no data-path collective)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alignet_b200 import shard, synth
from oracle import oracle


def test_batch_slices_partition_the_batch():
    for B in (0, 1, 2, 7, 64, 65):
        for w in (1, 2, 3, 8):
            s = shard.all_slices(B, w)
            assert s[0][0] == 0 and s[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(s, s[1:]))
            sizes = [hi - lo for lo, hi in s]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.batch_slice(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)  # every rank builds the same global batch, then keeps its slice
        vol = synth.volume("v1", B, 6, 7, 8, 2, rng)
        grid = synth.grid("g4", B, 5, 6, 9, rng)
        go = synth.grad_output(B, 5, 6, 9, 2, rng)
        v, g, o = shard.shard([vol, grid, go], world, rank)
        out = oracle.sampler_forward(np.ascontiguousarray(v), np.ascontiguousarray(g))
        gv, gg = oracle.sampler_backward(np.ascontiguousarray(v), np.ascontiguousarray(g), np.ascontiguousarray(o))
        full_out = shard.gather_batch(torch.from_numpy(out), B)
        full_gv = shard.gather_batch(torch.from_numpy(gv), B)
        full_gg = shard.gather_batch(torch.from_numpy(gg), B)
        slowest = shard.max_over_ranks(float(rank + 1))
        if rank == 0:
            ref_out = oracle.sampler_forward(vol, grid)
            ref_gv, ref_gg = oracle.sampler_backward(vol, grid, go)
            q.put((np.array_equal(full_out.numpy().view(np.uint32), ref_out.view(np.uint32)),
                   np.array_equal(full_gv.numpy().view(np.uint32), ref_gv.view(np.uint32)),
                   np.array_equal(full_gg.numpy().view(np.uint32), ref_gg.view(np.uint32)),
                   slowest))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [5, 8])
def test_two_gloo_ranks_reproduce_the_unsharded_result(B):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == (True, True, True, float(world))
