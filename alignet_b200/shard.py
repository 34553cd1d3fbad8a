"""Batch sharding of the warp path across ranks (one process per GPU).

The path partitions by sample: batch is the outermost loop of every reference function
(stn3dtorch/generic/BilinearSamplerVolumetric.c:482,641; Cumsum.lua:17-19;
stn3dtorch/VolumetricGridGenerator.lua:58-60) and the gradVol scatter stays inside sample b
(...c:676,686 add stride_b * b).  So ranks own contiguous batch slices and the data path needs no
collective; torch.distributed is used only to agree on timing (max over ranks) and, for callers that
want the full result on every rank, to all-gather the slices.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def batch_slice(batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the samples rank `rank` owns: contiguous, sizes differ by at most one, the first
    `batch % world_size` ranks take the extra sample."""
    if world_size < 1 or not (0 <= rank < world_size) or batch < 0:
        raise ValueError(f"bad sharding request: batch={batch} world_size={world_size} rank={rank}")
    base, extra = divmod(batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_slices(batch: int, world_size: int) -> List[Tuple[int, int]]:
    return [batch_slice(batch, world_size, r) for r in range(world_size)]


def shard(tensors: Sequence, world_size: int, rank: int):
    """Slice every tensor of a call (all B-leading) to this rank's samples."""
    lo, hi = batch_slice(tensors[0].shape[0], world_size, rank)
    return [t[lo:hi] for t in tensors]


def gather_batch(local, batch: int, group=None):
    """All-gather variable-length batch slices back into a B-leading tensor (every rank gets it).
    Works on whatever backend the group uses (NCCL on GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [hi - lo for lo, hi in all_slices(batch, world)]
    big = max(sizes) if sizes else 0
    pad = torch.zeros((big,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


def max_over_ranks(value: float, device=None, group=None) -> float:
    """The timing rule of the bench: a multi-GPU number is the slowest rank's."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
