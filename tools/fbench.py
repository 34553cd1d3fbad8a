#!/usr/bin/env python
"""tools/fbench.py -- developer timing of the fused / row kernels against the un-fused module chains they replace
(not the contract bench; bench.py reports the same comparisons under `fused`).

    python tools/fbench.py [--size 64 --batch 64 --csz 8 --iters 10] [--what row,affine,cage,cf]

Each variant: `iters` launches back to back between two CUDA events on the launch stream after 3 warm-ups; the
working set exceeds L2 at the default size (pass --flush for small sizes: 256 MiB write before every launch).
"""
import argparse, ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from alignet_b200 import _abi, synth
from alignet_b200 import nn as vnn

PEAK = 6538.3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=64); ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--csz", type=int, default=8); ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--what", default="row,affine,cage,cf"); ap.add_argument("--flush", action="store_true")
    ap.add_argument("--channels", type=int, default=4, help="channels of the channel-first case")
    a = ap.parse_args()
    N, B, csz = a.size, a.batch, a.csz
    V = B * N ** 3
    lib = _abi.lib(); d = vnn._desc
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if a.flush else None

    def timed(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        if flush is not None:
            ts = []
            for _ in range(a.iters):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return float(np.median(ts)) * 1e3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.iters * 1e3

    def timed_graph(fn):
        """the call captured once in a CUDA graph and replayed: device time without the Python wrapper (~0.1 ms per call)"""
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2): fn()
        torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters): g.replay()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.iters * 1e3

    def report(name, us, alg_bytes=None, note=""):
        frac = f" {alg_bytes / us / 1e3:7.0f} GB/s alg ({alg_bytes / us / 1e3 / PEAK:5.1%})" if alg_bytes else ""
        print(f"{name:58s} {us:8.1f} us  {V / us / 1e3:6.1f} Gvox/s{frac}  {note}", flush=True)

    rng = np.random.default_rng(7)
    nb = min(B, 8)
    rep = lambda x: torch.from_numpy(x).cuda().repeat(B // nb, *([1] * (x.ndim - 1))).contiguous()
    vol = rep(synth.volume("v2", nb, N, N, N, 1, rng))
    go = torch.randn((B, N, N, N, 1), device="cuda")
    out = torch.empty_like(go); gv = torch.empty_like(vol)
    what = a.what.split(",")

    if "row" in what:
        for gk in ("g1", "g2"):
            grid = rep(synth.grid(gk, nb, N, N, N, rng)); gg = torch.empty_like(grid)
            for algo, nm in ((0, "packed"), (5, "row")):
                fl = algo << 8
                f = timed(lambda: _abi.check(lib.volstn_b200_sampler_forward(0, 4, d(vol), d(grid), d(out), fl, st), "f"))
                b = timed(lambda: _abi.check(lib.volstn_b200_sampler_backward(0, 4, d(vol), d(grid), d(gv), d(gg), d(go), fl | _abi.BWD_ZERO_GRADVOL, st), "b"))
                bo = timed(lambda: _abi.check(lib.volstn_b200_sampler_backward(0, 4, d(vol), d(grid), d(gv), d(gg), d(go), fl | _abi.BWD_ONLY_GRID, st), "b"))
                report(f"sampler {gk} {nm} fwd", f, 24 * V); report(f"sampler {gk} {nm} bwd+memset", b, 44 * V)
                report(f"sampler {gk} {nm} bwd onlyGrid", bo, 40 * V)
            del grid, gg

    if "affine" in what:
        T = torch.eye(4, device="cuda").repeat(B, 1, 1) + 0.02 * torch.randn(B, 4, 4, device="cuda")
        gen, smp = vnn.VolumetricGridGenerator(N, N, N), vnn.BilinearSamplerVolumetric()
        fus = vnn.AffineSamplerVolumetric(N, N, N)
        grid = gen.updateOutput(T)
        f_gen = timed(lambda: gen.updateOutput(T)); f_s = timed(lambda: smp.updateOutput([vol, grid]))
        f_fus = timed(lambda: fus.updateOutput([vol, T]))
        report("affine chain fwd: gridgen + sampler", f_gen + f_s, 40 * V, f"({f_gen:.1f} + {f_s:.1f})")
        report("affine FUSED fwd", f_fus, 8 * V, f"x{(f_gen + f_s) / f_fus:.2f}")
        fus_row = vnn.AffineSamplerVolumetric(N, N, N, algo=_abi.ALGO_ROW)
        report("affine FUSED fwd (row kernel)", timed(lambda: fus_row.updateOutput([vol, T])), 8 * V)
        _, gg = smp.updateGradInput([vol, grid], go)
        b_s = timed(lambda: smp.updateGradInput([vol, grid], go)); b_gen = timed(lambda: gen.updateGradInput(T, gg))
        b_fus = timed(lambda: fus.updateGradInput([vol, T], go))
        report("affine chain bwd: sampler bwd + gridgen bwd", b_s + b_gen, 60 * V, f"({b_s:.1f} + {b_gen:.1f})")
        report("affine FUSED bwd (gradVol + gradT)", b_fus, 12 * V, f"x{(b_s + b_gen) / b_fus:.2f}")
        del grid, gg

    if "cagefast" in what:
        # the VOLSTN_FAST_ARITH lines only (brick kernels): quick A/B of library variants
        cage = torch.from_numpy(synth.alignet_cage(nb, csz, rng).astype(np.float32)).cuda().repeat(B // nb, 1, 1, 1, 1).contiguous()
        fast, fast_full = vnn.CageWarpVolumetric(onlyGrid=True, fastArith=True), vnn.CageWarpVolumetric(fastArith=True)
        tgt = rep(synth.volume("v2", nb, N, N, N, 1, rng))
        report("cage FUSED fwd, fast arithmetic", timed(lambda: fast.updateOutput([vol, cage])), 8 * V)
        report("cage FUSED bwd (gradCage only), fast arithmetic", timed(lambda: fast.updateGradInput([vol, cage], go)), 8 * V)
        report("cage FUSED bwd (gradSource + gradCage), fast arithmetic", timed(lambda: fast_full.updateGradInput([vol, cage], go)), 12 * V)
        stp = vnn.CageWarpSmoothL1Volumetric(fastArith=True)
        report("cage STEP (fwd + SmoothL1 + IoU + bwd to the cage), fast arithmetic", timed(lambda: stp.forwardBackward([vol, cage], tgt)), 8 * V)
        idc = torch.from_numpy(synth.alignet_cage(nb, csz, rng, jitter=0.0).astype(np.float32)).cuda().repeat(B // nb, 1, 1, 1, 1).contiguous()
        report("cage STEP, fast arithmetic, IDENTITY cage", timed(lambda: stp.forwardBackward([vol, idc], tgt)), 8 * V)
        report("GRAPH cage fwd, fast arithmetic", timed_graph(lambda: fast.updateOutput([vol, cage])), 8 * V)
        report("GRAPH cage bwd (gradCage only), fast arithmetic", timed_graph(lambda: fast.updateGradInput([vol, cage], go)), 8 * V)
        report("GRAPH cage STEP, fast arithmetic", timed_graph(lambda: stp.forwardBackward([vol, cage], tgt)), 8 * V)
        report("GRAPH cage STEP, fast arithmetic, IDENTITY cage", timed_graph(lambda: stp.forwardBackward([vol, idc], tgt)), 8 * V)
        report("GRAPH cage fwd, fast arithmetic, IDENTITY cage", timed_graph(lambda: fast.updateOutput([vol, idc])), 8 * V)
        for jit in (0.02, 0.05, 0.1):
            sc = torch.from_numpy(synth.alignet_cage(nb, csz, rng, jitter=jit).astype(np.float32)).cuda().repeat(B // nb, 1, 1, 1, 1).contiguous()
            report(f"cage STEP, fast arithmetic, cage jitter {jit}", timed(lambda: stp.forwardBackward([vol, sc], tgt)), 8 * V)

    if "cage" in what:
        cage = torch.from_numpy(synth.alignet_cage(nb, csz, rng).astype(np.float32)).cuda().repeat(B // nb, 1, 1, 1, 1).contiguous()
        cage_vol = torch.ones((B, csz, csz, csz, 4), device="cuda"); cage_vol[..., :3] = cage.permute(0, 2, 3, 4, 1)
        field_i = vnn.VolumetricGridGenerator(N, N, N).updateOutput(torch.eye(4, device="cuda").repeat(B, 1, 1))
        up, warp = vnn.BilinearSamplerVolumetric(onlyVolume=True), vnn.BilinearSamplerVolumetric(onlyGrid=True)
        warp_full = vnn.BilinearSamplerVolumetric()
        fus, fus_full = vnn.CageWarpVolumetric(onlyGrid=True), vnn.CageWarpVolumetric()
        field = up.updateOutput([cage_vol, field_i])
        f_up = timed(lambda: up.updateOutput([cage_vol, field_i])); f_w = timed(lambda: warp.updateOutput([vol, field]))
        f_fus = timed(lambda: fus.updateOutput([vol, cage]))
        report("cage chain fwd: up-sample(C=4) + warp", f_up + f_w, 56 * V, f"({f_up:.1f} + {f_w:.1f}; fieldI constant, not timed)")
        report("cage FUSED fwd", f_fus, 8 * V, f"x{(f_up + f_w) / f_fus:.2f}")
        _, gfield = warp.updateGradInput([vol, field], go)
        b_w = timed(lambda: warp.updateGradInput([vol, field], go)); b_up = timed(lambda: up.updateGradInput([cage_vol, field_i], gfield))
        b_fus = timed(lambda: fus.updateGradInput([vol, cage], go))
        report("cage chain bwd (onlyGrid warp + onlyVolume up-sample)", b_w + b_up, 72 * V, f"({b_w:.1f} + {b_up:.1f})")
        report("cage FUSED bwd (gradCage only)", b_fus, 8 * V, f"x{(b_w + b_up) / b_fus:.2f}")
        b_wf = timed(lambda: warp_full.updateGradInput([vol, field], go))
        b_ff = timed(lambda: fus_full.updateGradInput([vol, cage], go))
        report("cage chain bwd (full warp bwd + up-sample bwd)", b_wf + b_up, 76 * V, f"({b_wf:.1f} + {b_up:.1f})")
        report("cage FUSED bwd (gradSource + gradCage)", b_ff, 12 * V, f"x{(b_wf + b_up) / b_ff:.2f}")
        # VOLSTN_FAST_ARITH: separable cage up-sample, FMA sums, factored gradGrid (tolerance mode)
        fast, fast_full = vnn.CageWarpVolumetric(onlyGrid=True, fastArith=True), vnn.CageWarpVolumetric(fastArith=True)
        report("cage FUSED fwd, fast arithmetic", timed(lambda: fast.updateOutput([vol, cage])), 8 * V, f"x{f_fus / timed(lambda: fast.updateOutput([vol, cage])):.2f} over exact")
        report("cage FUSED bwd (gradCage only), fast arithmetic", timed(lambda: fast.updateGradInput([vol, cage], go)), 8 * V)
        report("cage FUSED bwd (gradSource + gradCage), fast arithmetic", timed(lambda: fast_full.updateGradInput([vol, cage], go)), 12 * V)
        tgt = rep(synth.volume("v2", nb, N, N, N, 1, rng))
        for fa in (False, True):
            stp = vnn.CageWarpSmoothL1Volumetric(fastArith=fa)
            report(f"cage STEP (fwd + SmoothL1 + IoU + bwd to the cage){', fast arithmetic' if fa else ''}", timed(lambda: stp.forwardBackward([vol, cage], tgt)), 8 * V)
        del field, field_i, gfield

    if "cf1" in what:
        # channel-first field (B x 4 x D x H x W, what a conv net emits) with a single-channel volume: the packed kernels
        # with planar component loads against the native layout and against the row kernels on the same views
        for gk in ("g1", "g2"):
            grid = rep(synth.grid(gk, nb, N, N, N, rng))
            gridp = grid.permute(0, 4, 1, 2, 3).contiguous()
            volp = vol.permute(0, 4, 1, 2, 3).contiguous(); gop = go.permute(0, 4, 1, 2, 3).contiguous()
            for nm, m, args, garg in (("native layout (packed)", vnn.BilinearSamplerVolumetric(), [vol, grid], go),
                                      ("channel-first, packed+planar grid", vnn.BilinearSamplerVolumetric(layout="BCDHW"), [volp, gridp], gop),
                                      ("channel-first, row kernels", vnn.BilinearSamplerVolumetric(layout="BCDHW", algo=_abi.ALGO_ROW), [volp, gridp], gop)):
                f = timed(lambda: m.updateOutput(args)); b = timed(lambda: m.updateGradInput(args, garg))
                report(f"cf1 {gk} {nm} fwd", f, 24 * V); report(f"cf1 {gk} {nm} bwd+memset", b, 44 * V)
            del grid, gridp

    if "cf" in what:
        C = a.channels
        Bc = max(1, B // C)
        Vc = Bc * N ** 3
        volc = torch.randn((Bc, C, N, N, N), device="cuda"); goc = torch.randn((Bc, C, N, N, N), device="cuda")
        gridc = rep(synth.grid("g2", nb, N, N, N, rng))[:Bc].permute(0, 4, 1, 2, 3).contiguous()      # B x 4 x D x H x W
        cfm, clm = vnn.BilinearSamplerVolumetric(layout="BCDHW"), vnn.BilinearSamplerVolumetric()
        f_cf = timed(lambda: cfm.updateOutput([volc, gridc])); b_cf = timed(lambda: cfm.updateGradInput([volc, gridc], goc))

        def chain_f():
            o = clm.updateOutput([volc.permute(0, 2, 3, 4, 1).contiguous(), gridc.permute(0, 2, 3, 4, 1).contiguous()])
            return o.permute(0, 4, 1, 2, 3).contiguous()

        def chain_b():
            gvv, ggg = clm.updateGradInput([volc.permute(0, 2, 3, 4, 1).contiguous(), gridc.permute(0, 2, 3, 4, 1).contiguous()],
                                           goc.permute(0, 2, 3, 4, 1).contiguous())
            return gvv.permute(0, 4, 1, 2, 3).contiguous(), ggg.permute(0, 4, 1, 2, 3).contiguous()

        f_ch, b_ch = timed(chain_f), timed(chain_b)
        s = V / Vc
        print(f"channel-first C={C} B={Bc} (g2):", flush=True)
        report("  transpose copies + packed sampler fwd", f_ch * s); report("  in-place channel-first fwd", f_cf * s, (16 + 8 * C) * V, f"x{f_ch / f_cf:.2f}")
        report("  transpose copies + packed sampler bwd", b_ch * s); report("  in-place channel-first bwd", b_cf * s, (32 + 12 * C) * V, f"x{b_ch / b_cf:.2f}")


if __name__ == "__main__":
    main()
