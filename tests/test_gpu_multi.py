"""Batch sharding on real GPUs: every visible device warps its own slice through the C ABI (device
pointers, that device's stream); the concatenation must equal the single-GPU result.  Runs with one
GPU too (then it only checks that slicing on one device is exact)."""
import numpy as np
import pytest
import torch

from alignet_b200 import nn as vnn, shard, synth

pytestmark = pytest.mark.gpu


def test_sharded_result_equals_single_device():
    ndev = torch.cuda.device_count()
    world = max(2, min(ndev, 4))
    rng = np.random.default_rng(3)
    B, N = 6, 24
    vol = torch.from_numpy(synth.volume("v1", B, N, N, N, 1, rng))
    grid = torch.from_numpy(synth.grid("g2", B, N, N, N, rng, csz=4))
    go = torch.from_numpy(synth.grad_output(B, N, N, N, 1, rng))
    m = vnn.BilinearSamplerVolumetric()
    ref_out = m.updateOutput([vol.cuda(0), grid.cuda(0)]).cpu()
    ref_gv, ref_gg = [t.cpu() for t in m.updateGradInput([vol.cuda(0), grid.cuda(0)], go.cuda(0))]
    outs, gvs, ggs = [], [], []
    for r in range(world):
        dev = torch.device("cuda", r % ndev)
        v, g, o = [t.to(dev) for t in shard.shard([vol, grid, go], world, r)]
        mr = vnn.BilinearSamplerVolumetric()
        outs.append(mr.updateOutput([v.contiguous(), g.contiguous()]).cpu())
        gv, gg = mr.updateGradInput([v.contiguous(), g.contiguous()], o.contiguous())
        gvs.append(gv.cpu()); ggs.append(gg.cpu())
    assert torch.equal(torch.cat(outs), ref_out)
    assert torch.equal(torch.cat(ggs), ref_gg)
    gv = torch.cat(gvs)
    assert float((gv - ref_gv).abs().max()) <= 1e-4 * float(ref_gv.abs().max())


def test_sharded_fused_step_equals_single_device():
    """the one-kernel training step shards by sample like everything else: losses add, cage gradients concatenate"""
    ndev = torch.cuda.device_count()
    world = max(2, min(ndev, 4))
    rng = np.random.default_rng(4)
    B, N = 6, 20
    src = torch.from_numpy(synth.volume("v2", B, N, N, N, 1, rng))
    tgt = torch.from_numpy(synth.volume("v2", B, N, N, N, 1, rng))
    cage = torch.from_numpy(synth.alignet_cage(B, 4, rng).astype(np.float32))
    ref = vnn.CageWarpSmoothL1Volumetric(sizeAverage=False)
    ref_loss, (_, ref_gc) = ref.forwardBackward([src.cuda(0), cage.cuda(0)], tgt.cuda(0))
    loss, gcs, counts = 0.0, [], []
    for r in range(world):
        dev = torch.device("cuda", r % ndev)
        s, c, t = [x.to(dev).contiguous() for x in shard.shard([src, cage, tgt], world, r)]
        m = vnn.CageWarpSmoothL1Volumetric(sizeAverage=False)
        l, (_, gc) = m.forwardBackward([s, c], t)
        loss += float(l)
        gcs.append(gc.cpu()); counts.append(m.iouCounts.cpu())
    assert abs(loss - float(ref_loss)) <= 1e-9 * max(1.0, float(ref_loss))
    assert torch.equal(torch.cat(counts), ref.iouCounts.cpu())
    gc = torch.cat(gcs)
    assert float((gc - ref_gc.cpu()).abs().max()) <= 1e-4 * float(ref_gc.abs().max())
