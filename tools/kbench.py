#!/usr/bin/env python
"""tools/kbench.py -- developer timing of kernel variants (not the contract bench; see bench.py).

    python tools/kbench.py [--size 64 --batch 64 --channels 1] [--grids g1,g2,g3] [--algos 1,2] [--iters 10]
Times volstn_b200_sampler_forward / _backward (device pointers, C ABI) per variant: `iters` launches
back to back between two CUDA events on the launch stream, after 3 warm-ups; the working set must
exceed L2 (or pass --flush).  Also checks each variant's gradVol / out against the direct variant.
"""
import argparse, ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from alignet_b200 import _abi, synth
from alignet_b200 import nn as vnn

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=64); ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--channels", type=int, default=1); ap.add_argument("--grids", default="g1,g2,g3")
    ap.add_argument("--algos", default="1,2"); ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--flags", default="0", help="extra flag bits (hex ok), comma list = more variants")
    a = ap.parse_args()
    N, B, C = a.size, a.batch, a.channels
    lib = _abi.lib(); d = vnn._desc
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    peak = 6538.3
    V = B * N ** 3
    for gk in a.grids.split(","):
        rng = np.random.default_rng(7)
        nb = min(B, 8)
        vol = torch.from_numpy(synth.volume("v2", nb, N, N, N, C, rng)).cuda().repeat(B // nb, 1, 1, 1, 1).contiguous()
        grid = torch.from_numpy(synth.grid(gk, nb, N, N, N, rng)).cuda().repeat(B // nb, 1, 1, 1, 1).contiguous()
        go = torch.randn((B, N, N, N, C), device="cuda")
        out = torch.empty((B, N, N, N, C), device="cuda"); gv = torch.empty_like(vol); gg = torch.empty_like(grid)
        ref = {}
        for algo in [int(x) for x in a.algos.split(",")]:
          for extra in [int(x, 0) for x in a.flags.split(",")]:
            flags = (algo << 8) | extra
            def fwd():
                _abi.check(lib.volstn_b200_sampler_forward(0, 4, d(vol), d(grid), d(out), flags, st), "fwd")
            def bwd():
                _abi.check(lib.volstn_b200_sampler_backward(0, 4, d(vol), d(grid), d(gv), d(gg), d(go), flags | _abi.BWD_ZERO_GRADVOL, st), "bwd")
            res = {}
            for name, fn in (("fwd", fwd), ("bwd", bwd)):
                for _ in range(3): fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.iters): fn()
                e1.record(); torch.cuda.synchronize()
                res[name] = e0.elapsed_time(e1) / a.iters * 1e3
            chk = ""
            if not ref:
                ref = {"out": out.clone(), "gv": gv.clone(), "gg": gg.clone()}
            else:
                chk = "out_eq=%s gg_eq=%s gv_rel=%.2e" % (torch.equal(out, ref["out"]), torch.equal(gg, ref["gg"]),
                      float((gv - ref["gv"]).abs().max() / ref["gv"].abs().max()))
            fb, bb = (16 + 8 * C) * V, (32 + 12 * C) * V
            print(f"{gk} N={N} B={B} C={C} algo={algo} extra={extra:#x}: fwd {res['fwd']:7.1f} us ({fb/res['fwd']/1e3/peak:5.1%})  "
                  f"bwd+memset {res['bwd']:7.1f} us ({bb/res['bwd']/1e3/peak:5.1%})  step {V/(res['fwd']+res['bwd'])/1e3:6.1f} Gvox/s "
                  f"({(fb+bb)/(res['fwd']+res['bwd'])/1e3/peak:5.1%})  {chk}", flush=True)
        del vol, grid, go, out, gv, gg
if __name__ == "__main__":
    main()
