#!/usr/bin/env python
"""tools/train_step.py -- BASELINE.json config 3: an ALIGNet-3D training step at 64^3, global batch 64,
random-init weights, batch-sharded over the ranks (SURVEY.md section 3.3 with the bracketed 3-D lift, section 8e).

    python tools/train_step.py [--steps 10]                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/train_step.py

The step is what train.lua:106-141 does per batch: CNN on the stacked source/target pair -> cage deltas
-> Abs -> nn.Cumsum -> +beta -> cage up-sample (sampler, volume = cage, grid = identity field from
nn.VolumetricGridGenerator) -> source warp (sampler) -> SmoothL1 -> backward -> Adam.  The CNN is
the 3-D lift of models/alignet_2d.lua:33-57 in plain PyTorch (library code, NOT part of this repo's
product); Cumsum, the grid generator and both samplers are this repo's kernels through the C ABI.
The only collective is one NCCL all-reduce of the flattened CNN gradients (the warp path shards by
sample and needs none).  Prints one JSON line: step time, samples/s, and the warp path's share.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import torch.nn as tnn

from alignet_b200 import autograd as vag, nn as vnn, shard, synth, _abi


class WarpNet(tnn.Module):
    """3-D lift of warpF + dNet (models/alignet_2d.lua:33-120), csz = 4 for 64^3 inputs."""

    def __init__(self, csz=4):
        super().__init__()
        self.csz = csz
        self.features = tnn.Sequential(
            tnn.MaxPool3d(2), tnn.Conv3d(2, 20, 5), tnn.ReLU(True),
            tnn.MaxPool3d(2), tnn.Conv3d(20, 20, 5), tnn.ReLU(True),
            tnn.MaxPool3d(2), tnn.Conv3d(20, 20, 2), tnn.ReLU(True),
            tnn.Conv3d(20, 20, 1), tnn.ReLU(True))                      # 64 -> 32 -> 28 -> 14 -> 10 -> 5 -> 4
        self.fc = tnn.Sequential(tnn.Linear(20 * csz ** 3, 20), tnn.ReLU(True))
        self.out = tnn.Linear(20, 3 * csz ** 3)
        self.delta = 2.0 / (csz - 1)
        with torch.no_grad():                                           # identity warp at init (lua:77-80,118-119)
            self.out.weight.zero_()
            self.out.bias.fill_(self.delta)
        self.beta = tnn.Parameter(torch.tensor([-1.0 - self.delta]))    # learn_beta (lua:104-108)
        self.cumsum = vnn.Cumsum()

    def forward(self, pair):
        B = pair.shape[0]
        d = self.out(self.fc(self.features(pair).reshape(B, -1)))
        d = d.abs().view(B, 3, self.csz, self.csz, self.csz)            # range_fct = abs (lua:84-92)
        return vag.cumsum(d.contiguous(), self.cumsum) + self.beta      # nn.Cumsum -> + beta (lua:102-112)


def run(steps=10, warmup=3, global_batch=64, size=64, fused=2, rank=0, world=1, graph=False):
    """The training step on an ALREADY initialised process group (bench.py --config train and the bench line's `train`
    section call this; main() below is the stand-alone launcher).  Returns the result dict on every rank."""
    N, csz = size, 4
    if os.environ.get("VOLSTN_TRAIN_CUDNN_BENCHMARK", "1") != "0":
        torch.backends.cudnn.benchmark = True       # let cuDNN time its Conv3d algorithms once (library code, not product)
    lo, hi = shard.batch_slice(global_batch, world, rank)
    B = hi - lo
    torch.manual_seed(0)                                                # identical weights on every rank
    net = WarpNet(csz).cuda()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=bool(graph))
    rng = np.random.default_rng(100 + rank)
    src = torch.from_numpy(synth.volume("v2", B, N, N, N, 1, rng)).cuda()          # B x D x H x W x 1
    tgt = torch.from_numpy(synth.volume("v2", B, N, N, N, 1, rng)).cuda()
    pair = torch.stack([src[..., 0], tgt[..., 0]], dim=1).contiguous()              # B x 2 x D x H x W
    field_i = None
    if fused == 0:
        gen = vnn.VolumetricGridGenerator(N, N, N)
        field_i = gen.updateOutput(torch.eye(4, device="cuda").repeat(B, 1, 1))     # constant, like fieldI (dataset_2d.lua:128)
    up, warp = vnn.BilinearSamplerVolumetric(), vnn.BilinearSamplerVolumetric()
    fwarp, fstep = vnn.CageWarpVolumetric(), vnn.CageWarpSmoothL1Volumetric(fastArith=(fused == 3))
    params = [p for p in net.parameters()]
    crit = tnn.SmoothL1Loss()
    ones = torch.ones((B, csz, csz, csz, 1), device="cuda")
    # ONE flat gradient buffer, every parameter's .grad a view into it: the all-reduce needs no flatten / un-flatten copies
    flat = torch.zeros(sum(p.numel() for p in params), device="cuda")
    o = 0
    for p in params:
        p.grad = flat[o:o + p.numel()].view_as(p)
        o += p.numel()
    ev = {k: [torch.cuda.Event(enable_timing=True) for _ in range(2)] for k in ("cnn_fwd", "warp", "cnn_bwd", "allreduce", "adam")}

    def mark(timed, key, i):
        if timed:
            ev[key][i].record()

    def step(timed=False):
        flat.zero_()                                                               # = opt.zero_grad(), one memset
        mark(timed, "cnn_fwd", 0)
        cage = net(pair)                                                            # B x 3 x csz^3
        mark(timed, "cnn_fwd", 1)
        mark(timed, "warp", 0)
        if fused >= 2:                                                              # warp + criterion + both backwards: one kernel
            loss = vag.cage_warp_smooth_l1(src, cage.contiguous(), tgt, fstep)
        else:
            if fused == 1:
                est = vag.cage_warp_volumetric(src, cage.contiguous(), fwarp)       # cage up-sample + source warp: one kernel
            else:
                cage_vol = torch.cat([cage.permute(0, 2, 3, 4, 1), ones], dim=-1).contiguous()   # channels-last, 4th coord
                field = vag.bilinear_sampler_volumetric(cage_vol, field_i, up)      # cage up-sample -> B x N^3 x 4
                est = vag.bilinear_sampler_volumetric(src, field, warp)             # source warp
            loss = crit(est, tgt)
        mark(timed, "warp", 1)
        mark(timed, "cnn_bwd", 0)
        loss.backward()
        mark(timed, "cnn_bwd", 1)
        mark(timed, "allreduce", 0)
        if world > 1:                                                               # the one collective of the step
            dist.all_reduce(flat)
            flat.div_(world)
        mark(timed, "allreduce", 1)
        mark(timed, "adam", 0)
        opt.step()
        mark(timed, "adam", 1)
        return loss

    if graph:
        # Every step of a graph run -- the warm-ups included -- is issued on a side stream: autograd binds a parameter's
        # gradient accumulator to the stream that was current when the parameter was first used, and an accumulator bound
        # to the legacy default stream cannot take part in a capture.
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        torch.cuda.set_stream(side)
    for _ in range(max(warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    n0 = _abi.launch_count()
    step()
    launches = _abi.launch_count() - n0
    run_step = step
    if graph:
        # the whole step -- CNN forward, warp head, both backwards, the all-reduce, Adam -- as ONE CUDA graph: at 8 samples per
        # GPU the eager step is bound by the ~150 kernel launches of the CNN, not by the GPU
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            static_loss = step()
        torch.cuda.synchronize()

        def run_step():
            g.replay()
            return static_loss
        for _ in range(2):
            run_step()
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = run_step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        dist.barrier()
        ms = shard.max_over_ranks(ms, device="cuda")
    # where the step goes: one more step with events around its stages (device time of each stage on this rank; for
    # fused >= 2 "warp" is forward + criterion + the warp's backward in one kernel and "cnn_bwd" the CNN's backward alone)
    parts = {k: 0.0 for k in ev}
    reps = 3
    step(); torch.cuda.synchronize()       # after a capture the first eager step re-runs cuDNN's algorithm search: not timed
    for _ in range(reps):
        step(True); torch.cuda.synchronize()
        for k in ev:
            parts[k] += ev[k][0].elapsed_time(ev[k][1]) / reps
    if world > 1:
        parts = {k: shard.max_over_ranks(v, device="cuda") for k, v in parts.items()}
        dist.barrier()
    nparam = sum(p.numel() for p in params)
    limiter = max(parts, key=parts.get)
    return {"config": "ALIGNet-3D training step (BASELINE.json config 3)", "size": N, "global_batch": global_batch,
            "n_gpus": world, "batch_per_gpu": B, "ms_per_step": ms, "samples_per_s": global_batch / (ms * 1e-3),
            "voxels_per_s": global_batch * N ** 3 / (ms * 1e-3), "loss": float(loss.detach()),
            "stage_ms": parts, "largest_stage": limiter, "fused": fused, "volstn_kernels_per_step": int(launches), "cuda_graph": bool(graph),
            "allreduce_floats": nparam if world > 1 else 0, "allreduce_us": parts["allreduce"] * 1e3, "scaling": "strong",
            "note": "CNN = plain PyTorch / cuDNN fp32 Conv3d (3-D lift of models/alignet_2d.lua:33-57; library code); Cumsum and the "
                    "warp head = this repo's kernels; one NCCL all-reduce of the CNN gradients per step (one flat buffer, .grad are views; "
                    "all-reduce + division timed as `allreduce`)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10); ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--global-batch", type=int, default=64); ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--fused", type=int, default=0, help="0: un-fused modules; 1: CageWarpVolumetric + torch criterion; 2: the one-kernel step; 3: the same in fast arithmetic (brick kernels)")
    ap.add_argument("--graph", type=int, default=0, help="1: replay the whole step as one CUDA graph (stage_ms stay eager)")
    a = ap.parse_args()
    out_fd = os.fdopen(os.dup(1), "w"); os.dup2(2, 1)                   # stdout = the JSON line only
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = run(a.steps, a.warmup, a.global_batch, a.size, a.fused, rank, world, bool(a.graph))
    if rank == 0:
        print(json.dumps(res), file=out_fd, flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
