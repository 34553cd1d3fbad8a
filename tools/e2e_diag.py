#!/usr/bin/env python
"""tools/e2e_diag.py -- where the host-tensor step's time goes (developer tool): synchronous calls one by one, one
asynchronous step, and the pipelined loop of bench.py's `e2e.variants`, with the enqueue time of every call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from alignet_b200 import nn as vnn, synth

N, B = 64, 64
rng = np.random.default_rng(0)
vol = torch.from_numpy(synth.volume("v2", B, N, N, N, 1, rng)).pin_memory()
grid = torch.from_numpy(synth.grid("g1", B, N, N, N, rng)).pin_memory()
go = torch.from_numpy(synth.grad_output(B, N, N, N, 1, rng)).pin_memory()


def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), sum(ts) / n


m = vnn.BilinearSamplerVolumetric()
print("sync fwd  ms (min, mean):", t(lambda: m.updateOutput([vol, grid])))
print("sync bwd  ms (min, mean):", t(lambda: m.updateGradInput([vol, grid], go)))
for reuse in (False, True):
    ma = [vnn.BilinearSamplerVolumetric(hostAsync=True, reuseForwardInputs=reuse) for _ in range(2)]

    def one_step():
        t0 = time.perf_counter(); ma[0].updateOutput([vol, grid]); t1 = time.perf_counter()
        ma[0].updateGradInput([vol, grid], go); t2 = time.perf_counter(); vnn.host_wait(); t3 = time.perf_counter()
        return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3

    one_step()
    ma[1].updateOutput([vol, grid]); ma[1].updateGradInput([vol, grid], go); vnn.host_wait()   # pinned outputs of the second set
    print(f"async reuse={reuse}: one step: enqueue fwd / enqueue bwd / wait ms:", [tuple(round(x, 2) for x in one_step()) for _ in range(3)])
    steps = 10
    tick = [None, None]
    t0 = time.perf_counter()
    enq = []
    for i in range(steps):
        k = i & 1
        if tick[k] is not None:
            vnn.host_wait(tick[k])
        a = time.perf_counter()
        ma[k].updateOutput([vol, grid]); ma[k].updateGradInput([vol, grid], go)
        enq.append((time.perf_counter() - a) * 1e3)
        tick[k] = vnn.host_ticket()
    vnn.host_wait()
    print(f"async reuse={reuse}: pipelined {steps} steps: {(time.perf_counter() - t0) / steps * 1e3:.2f} ms/step; enqueue ms per step:", [round(x, 2) for x in enq])
