#!/usr/bin/env python
"""bench.py -- sampled voxels/s of nn.BilinearSamplerVolumetric forward + backward at 64^3.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--size 64] [--batch 64] [--channels 1] [--grid g1] [--sweep] [--config sampler|train|contention]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = updateOutput + updateGradInput of the module (gradVol zero-fill included, as
BilinearSamplerVolumetric.lua:101-104 does it) over one batch of synthetic voxel volumes and the warp field a
random-init ALIGNet emits (`g1`, the identity field; the ALIGNet-like sheared field `g2` is timed in the same run
and reported as `value_g2`).  The batch shards over ranks with no collective (samples are independent); per-GPU
work is fixed, so scaling is "weak".

JSON line (rank 0): metric/value = whole-job sampled voxels/s with inputs resident in HBM;
`e2e` = the same through the HOST-buffer C ABI exactly as the drop-in Lua module calls it (pinned host tensors in,
host tensors out, every input re-uploaded by both calls, copies inside the timed region; `e2e.variants` holds the
opt-in modes); `roofline` = the dominant kernel (backward) against the measured HBM copy bandwidth; `cpu_baseline`
= the reference's own C code on this box's host cores; `train` / `contention` = BASELINE.json configs 3 and 5.
`--impl reference` times only that CPU implementation and prints the same line shape.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "sampled voxels/s fwd+bwd @64^3 (BilinearSamplerVolumetric)"
UNIT = "voxels/s"


def fwd_bytes(C, G=4):
    return 4 * G + 8 * C            # grid + volume + out            


def bwd_bytes(C, G=4):
    return 8 * G + 12 * C           # grid + gradOut + volume + gradGrid + gradVol


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(name):
    """dram bytes per launch of the named kernel from the committed ncu capture, or None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(name)
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------
# clocks during the timed region (NVML, in-process so that millisecond regions are still sampled)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.ok = False
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.ok:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *a):
        if self.ok:
            self._stop.set()
            self.t.join()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# the reference's CPU implementation, timed on the host cores
# ------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_reference_rate(N, C, grid_kind, batch, reps, threads, seed=0):
    """voxels/s of the reference CPU code (oracle/_ref, else the C port) for fwd+bwd on `batch`
    samples of the workload, batch-sliced over `threads` host threads; best of `reps`."""
    from alignet_b200 import synth
    from oracle import oracle
    impl = "ref" if oracle.have_ref() else "port"
    rng = np.random.default_rng(seed)
    vol = synth.volume("v2", batch, N, N, N, C, rng)
    grid = synth.grid(grid_kind, batch, N, N, N, rng)
    go = synth.grad_output(batch, N, N, N, C, rng)
    times = []
    bufs = {}
    oracle.sampler_forward_backward_threaded(vol, grid, go, 4, impl, threads, bufs)   # allocates the outputs once (lua: resize)
    for _ in range(reps):
        t0 = time.perf_counter()
        oracle.sampler_forward_backward_threaded(vol, grid, go, 4, impl, threads, bufs)
        times.append(time.perf_counter() - t0)
    best = min(times)
    return batch * N ** 3 / best, statistics.mean(times), ("reference" if impl == "ref" else "port")


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm; the others exit 0 without work
    threads = host_threads()
    N, C = a.size, a.channels
    batch = a.batch  # the whole per-GPU batch of the workload (64 samples of 64^3: ~0.1 s per step on 16 threads)
    from alignet_b200 import synth
    from oracle import oracle
    impl = "ref" if oracle.have_ref() else "port"
    rng = np.random.default_rng(0)
    vol = synth.volume("v2", batch, N, N, N, C, rng)
    grid = synth.grid(a.grid, batch, N, N, N, rng)
    go = synth.grad_output(batch, N, N, N, C, rng)
    bufs = {}
    for _ in range(max(a.warmup, 1)):
        oracle.sampler_forward_backward_threaded(vol, grid, go, 4, impl, threads, bufs)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        oracle.sampler_forward_backward_threaded(vol, grid, go, 4, impl, threads, bufs)
    dt = time.perf_counter() - t0
    value = batch * N ** 3 * a.steps / dt
    kind = "reference" if impl == "ref" else "port"
    sample = (f"all {batch} samples of the {N}^3 C={C} {a.grid} workload per step (fwd + zero-fill + bwd), "
              f"batch-sliced over {threads} host threads ({'oracle/_ref = reference C compiled unmodified' if impl == 'ref' else 'oracle C port'}, gcc -O2 -ffp-contract=off)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(a), timed_region="K steps of the reference's C functions on host tensors, wall clock"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def workload_config(a, batch_override=None):
    N, C = a.size, a.channels
    B = batch_override if batch_override is not None else a.batch
    ws = B * N ** 3 * (2 * 16 + 4 * 4 * C) / 2 ** 20
    return {"workload": f"BilinearSamplerVolumetric updateOutput+updateGradInput, volume {N}^3 x C={C}, grid {N}^3 x 4 "
                        f"({a.grid}: {GRID_DESC.get(a.grid, a.grid)}), batch {B} per GPU",
            "size": N, "channels": C, "batch_per_gpu": B, "grid": a.grid,
            "l2": f"no flush: working set {ws:.0f} MiB per GPU exceeds the 126 MB L2" if ws > 252 else "L2 flushed (256 MiB write) between timed iterations",
            "parallelism": f"batch-sharded x{a.gpus}, no data-path collective"}


GRID_DESC = {"g1": "identity field = ALIGNet's random-init warp (models/alignet_2d.lua:118-119)", "g2": "ALIGNet-like monotone cage field up-sampled from 8^3",
             "g3": "uniform random U(-1,1)", "g4": "extreme/out-of-bounds mix", "g5": "random rotations"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist

    from alignet_b200 import _abi, shard, synth
    from alignet_b200 import nn as vnn

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- alignet_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    distributed = world > 1
    if distributed:
        # NCCL logs (even its version banner) go to stdout by default; stdout carries the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world if distributed else 1

    # pinned host buffers next to this rank's GPU: bind the thread to the GPU's local CPUs while they are allocated
    # and first touched (a no-op where the platform reports one node); restored before the CPU baseline runs
    full_affinity = os.sched_getaffinity(0)
    _abi.lib().volstn_b200_host_bind_thread(local)
    bound_cpus = len(os.sched_getaffinity(0))

    N, C, B = a.size, a.channels, a.batch
    V = B * N ** 3
    rng = np.random.default_rng(1000 + rank)
    vol_h = torch.from_numpy(synth.volume("v2", B, N, N, N, C, rng)).pin_memory()
    grid_h = torch.from_numpy(synth.grid(a.grid, B, N, N, N, rng)).pin_memory()
    go_h = torch.from_numpy(synth.grad_output(B, N, N, N, C, rng)).pin_memory()
    vol, grid, go = vol_h.cuda(), grid_h.cuda(), go_h.cuda()
    m = vnn.BilinearSamplerVolumetric(algo=a.algo << 8, vpt=a.vpt)
    ws_mib = B * N ** 3 * (2 * 16 + 4 * 4 * C) / 2 ** 20
    flush = None if ws_mib > 252 else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step():
        m.updateOutput([vol, grid])
        m.updateGradInput([vol, grid], go)

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        step()
    barrier()

    # The step is launch-bound from Python (three ~100 us kernels behind ctypes calls), so it is captured
    # once into a CUDA graph -- the module calls launch on the capturing stream through the C ABI -- and
    # the timed region replays that graph.
    graph = None
    launches_per_step = None
    if not a.no_graph:
        n0 = _abi.launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        launches_per_step = _abi.launch_count() - n0
        for _ in range(3):
            graph.replay()
        barrier()

    def run_step():
        if graph is not None:
            graph.replay()
        else:
            step()

    sampler = ClockSampler(local)
    with sampler:
        # ---- (1) headline: K steps, CUDA events on the launching stream, max over ranks ----
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _abi.launch_count()
        if flush is None:
            ev0.record()
            for _ in range(a.steps):
                run_step()
            ev1.record()
            barrier()
            total_ms = ev0.elapsed_time(ev1)
        else:
            total_ms = 0.0
            for _ in range(a.steps):
                flush.zero_()
                ev0.record()
                run_step()
                ev1.record()
                torch.cuda.synchronize()
                total_ms += ev0.elapsed_time(ev1)
            barrier()
        launches = (launches_per_step * a.steps) if graph is not None else (_abi.launch_count() - n0)
        if distributed:
            total_ms = shard.max_over_ranks(total_ms, device="cuda")
        ms_per_step = total_ms / a.steps
        value = V * n_gpus / (ms_per_step * 1e-3)

        # ---- (2) per-kernel durations: K launches of one kernel back to back between two events on
        # the launch stream (the queue stays full, so the CPU-side ctypes cost does not show) ----
        lib = _abi.lib()
        d = vnn._desc
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        out = torch.empty((B, N, N, N, C), device="cuda")
        gv, gg = torch.zeros_like(vol), torch.empty_like(grid)
        flags = (a.algo << 8) | _abi.VPT(a.vpt)
        dv, dg, do, dgv, dgg, dgo = d(vol), d(grid), d(out), d(gv), d(gg), d(go)

        def timed(fn):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            if flush is not None:
                ts = []
                for _ in range(a.steps):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                return statistics.mean(ts)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / a.steps

        fwd_ms = timed(lambda: _abi.check(lib.volstn_b200_sampler_forward(0, 4, dv, dg, do, flags, st), "fwd"))
        # accumulating into gradVol (no zero-fill flag) is the kernel alone, exactly what the reference's
        # C function does; the zero-fill is timed separately
        bwd_ms = timed(lambda: _abi.check(lib.volstn_b200_sampler_backward(0, 4, dv, dg, dgv, dgg, dgo, flags, st), "bwd"))
        zero_ms = timed(lambda: gv.zero_())

        # ---- (3) end to end through the HOST-buffer ABI: pinned host in, host out ----
        e2e = None
        if not a.no_e2e:
            e2e = run_e2e(a, torch, dist, vnn, _abi, shard, vol_h, grid_h, go_h, V, n_gpus, distributed, barrier, local)
    clocks = sampler.summary()

    peak, peak_src = measured_peak()
    bwd_alg = bwd_bytes(C) * V
    fwd_alg = fwd_bytes(C) * V
    bwd_gbs = bwd_alg / (bwd_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_sampler_bwd_packed (gradVol scatter-add + gradGrid)",
                "achieved": bwd_gbs, "peak": peak, "unit": "GB/s", "frac": bwd_gbs / peak,
                "traffic": ncu_traffic("bwd"), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bwd_alg, "avg_launch_ms": bwd_ms,
                "forward": {"achieved": fwd_alg / (fwd_ms * 1e-3) / 1e9, "frac": fwd_alg / (fwd_ms * 1e-3) / 1e9 / peak,
                            "algorithmic_bytes_per_launch": fwd_alg, "avg_launch_ms": fwd_ms, "traffic": ncu_traffic("fwd")},
                "gradvol_zero_fill_ms": zero_ms,
                "step": {"achieved": (fwd_alg + bwd_alg) / (ms_per_step * 1e-3) / 1e9,
                         "frac": (fwd_alg + bwd_alg) / (ms_per_step * 1e-3) / 1e9 / peak,
                         "note": "48+20C algorithmic B/voxel over the whole step incl. the gradVol zero-fill"}}

    os.sched_setaffinity(0, full_affinity)
    cpu = None
    if rank == 0 and not a.no_cpu:
        threads = host_threads()
        sample_b = B
        rate, mean_s, kind = cpu_reference_rate(N, C, a.grid, sample_b, reps=3, threads=threads)
        one_rate, _, _ = cpu_reference_rate(N, C, a.grid, min(B, 4), reps=1, threads=1)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"all {sample_b} samples of the same {N}^3 workload, fwd + zero-fill of both gradInputs + bwd into pre-allocated outputs, batch-sliced over {threads} host threads, best of 3",
               "single_core_value": one_rate}

    sweep = run_sweep(a, torch, vnn, synth, _abi) if (a.sweep and rank == 0) else None
    # context sections: they must never cost the contract line
    def guarded(fn):
        try:
            return fn(a, torch, vnn, synth, _abi, peak)
        except Exception as e:  # noqa: BLE001
            torch.cuda.synchronize()
            return {"error": repr(e)[:300]}

    aux = guarded(run_aux) if (rank == 0 and n_gpus == 1 and not a.no_aux) else None
    fused = guarded(run_fused) if (rank == 0 and n_gpus == 1 and not a.no_aux) else None

    # the same step on the other warp-field families (every rank, inputs resident, max over ranks): `g2`, the sheared
    # field a TRAINED ALIGNet emits, is the second headline (`value_g2`)
    fields = None
    if not a.no_fields:
        fields = {}
        peak_, _ = measured_peak()
        for gk in ("g1", "g2", "g3"):
            if gk == a.grid:
                continue
            rng2 = np.random.default_rng(7 + rank)
            nb = min(B, 8)
            g2 = torch.from_numpy(synth.grid(gk, nb, N, N, N, rng2)).cuda().repeat((B + nb - 1) // nb, 1, 1, 1, 1)[:B].contiguous()
            m2 = vnn.BilinearSamplerVolumetric(algo=a.algo << 8, vpt=a.vpt)
            for _ in range(3):
                m2.updateOutput([vol, g2]); m2.updateGradInput([vol, g2], go)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                m2.updateOutput([vol, g2]); m2.updateGradInput([vol, g2], go)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            if distributed:
                ms = shard.max_over_ranks(ms, device="cuda")
            fields[gk] = {"desc": GRID_DESC[gk], "ms_per_step": ms, "voxels_per_s": V * n_gpus / (ms * 1e-3),
                          "step_frac_of_hbm_peak": (fwd_alg + bwd_alg) / (ms * 1e-3) / 1e9 / peak_}
            if gk == "g2" and C == 1:
                # the opt-in TMA-staged backward (VOLSTN_ALGO_TMA) behind the automatic forward
                mt = vnn.BilinearSamplerVolumetric(algo=_abi.ALGO_TMA)
                for _ in range(3):
                    m2.updateOutput([vol, g2]); mt.updateGradInput([vol, g2], go)
                barrier()
                e0.record()
                for _ in range(reps):
                    m2.updateOutput([vol, g2]); mt.updateGradInput([vol, g2], go)
                e1.record(); torch.cuda.synchronize()
                ms_t = e0.elapsed_time(e1) / reps
                if distributed:
                    ms_t = shard.max_over_ranks(ms_t, device="cuda")
                fields[gk]["tma_backward"] = {"ms_per_step": ms_t, "voxels_per_s": V * n_gpus / (ms_t * 1e-3),
                                              "step_frac_of_hbm_peak": (fwd_alg + bwd_alg) / (ms_t * 1e-3) / 1e9 / peak_,
                                              "note": "automatic forward + VOLSTN_ALGO_TMA backward (opt-in: the identity field loses 40 % with it)"}
            del g2

    # BASELINE.json configs 3 and 5 (every rank takes part: config 3 has the path's one collective)
    extras = {}
    if not a.no_extra:
        for name, fn in (("train", run_train_config), ("contention", run_contention_config)):
            try:
                extras[name] = fn(a, torch, dist, rank, n_gpus, distributed)
            except Exception as e:  # noqa: BLE001  (context sections must never cost the contract line)
                torch.cuda.synchronize()
                extras[name] = {"error": repr(e)[:300]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(a),
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
        line["config"]["timed_region"] = ("CUDA graph of one step (forward, gradVol zero-fill, backward) replayed K times"
                                         if graph is not None else "K steps launched from Python")
        if fields is not None:
            line["other_fields"] = fields
            if "g2" in fields:
                line["value_g2"] = fields["g2"]["voxels_per_s"]
                line["value_g2_note"] = "the same step on the ALIGNet-like sheared field g2 (what a trained network emits), whole job, device-timed, automatic kernel choice"
                if "tma_backward" in fields["g2"]:
                    line["value_g2_best"] = max(fields["g2"]["voxels_per_s"], fields["g2"]["tma_backward"]["voxels_per_s"])
        line["config"]["host_buffers"] = f"pinned; thread bound to {bound_cpus} of {len(full_affinity)} CPUs next to the GPU while allocating (volstn_b200_host_bind_thread)"
        line.update(extras)
        if aux is not None:
            line["aux_kernels"] = aux
        if fused is not None:
            line["fused"] = fused
        if sweep is not None:
            line["sweep"] = sweep
        emit(line)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(a, torch, dist, vnn, _abi, shard, vol_h, grid_h, go_h, V, n_gpus, distributed, barrier, local):
    """The step through the HOST-buffer ABI, wall clock, copies inside the timed region, every step's result read on the
    host.  Headline = the calls lua/BilinearSamplerVolumetric.lua makes (synchronous; every input re-uploaded by both calls
    like the reference re-reads them, ...c:587-640; the gradVol zero-fill travels as a flag).  `variants`: the unmodified
    reference Lua file (flags = 0: host-side :zero() of both gradInputs, gradVol uploaded and accumulated into), and the
    opt-in modes (input reuse, asynchronous calls with two steps in flight)."""
    lib, d, esz = _abi.lib(), vnn._desc, 4
    nv, ng, no = vol_h.numel() * esz, grid_h.numel() * esz, go_h.numel() * esz
    e2e_steps = max(3, min(a.steps, 10))

    def timed_loop(step_fn, finish=None):
        for w in range(2):      # both buffer sets of the asynchronous variants allocate their pinned outputs here
            step_fn(w)
        if finish:
            finish()
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            step_fn(i)
        if finish:
            finish()
        barrier()
        sec = (time.perf_counter() - t0) / e2e_steps
        return shard.max_over_ranks(sec, device="cuda") if distributed else sec

    def sync_step(m):
        def f(i):
            o = m.updateOutput([vol_h, grid_h])
            g = m.updateGradInput([vol_h, grid_h], go_h)
            return float(o.view(-1)[0]) + float(g[1].view(-1)[0])  # results are host tensors already
        return f

    # headline: module defaults = the drop-in Lua module
    sec = timed_loop(sync_step(vnn.BilinearSamplerVolumetric(algo=a.algo << 8, vpt=a.vpt)))
    variants = {}
    # the reference's UNMODIFIED Lua file over the drop-in library: zero both gradInputs on the host, flags = 0
    ctx = vnn._host_ctx()
    out_h = torch.empty(tuple(go_h.shape)).pin_memory()
    gv_h, gg_h = torch.empty(tuple(vol_h.shape)).pin_memory(), torch.empty(tuple(grid_h.shape)).pin_memory()

    def ref_lua_step(i):
        _abi.check(lib.volstn_b200_host_sampler_forward(ctx, 0, 4, d(vol_h), d(grid_h), d(out_h), 0), "fwd")
        gv_h.zero_(); gg_h.zero_()                                     # lua:101-104
        _abi.check(lib.volstn_b200_host_sampler_backward(ctx, 0, 4, d(vol_h), d(grid_h), d(gv_h), d(gg_h), d(go_h), 0), "bwd")
        return float(out_h.view(-1)[0]) + float(gg_h.view(-1)[0])

    t = timed_loop(ref_lua_step)
    variants["reference_lua_file_flags0"] = {"ms_per_step": t * 1e3, "value": V * n_gpus / t,
                                             "h2d_bytes_per_step": 2 * (nv + ng) + no + nv, "d2h_bytes_per_step": no + nv + ng,
                                             "note": "stn3dtorch/BilinearSamplerVolumetric.lua unchanged: host :zero() of both gradInputs, gradVol uploaded (accumulate semantics)"}
    del out_h, gv_h, gg_h
    t = timed_loop(sync_step(vnn.BilinearSamplerVolumetric(algo=a.algo << 8, vpt=a.vpt, reuseForwardInputs=True)))
    variants["reuse_inputs"] = {"ms_per_step": t * 1e3, "value": V * n_gpus / t, "h2d_bytes_per_step": nv + ng + no, "d2h_bytes_per_step": no + nv + ng,
                                "note": "opt-in VOLSTN_HOST_KEEP/REUSE_INPUTS: updateGradInput re-uses updateOutput's device copies of volume and grid"}
    # asynchronous calls, two steps in flight (a module = a set of output buffers per step in flight)
    for reuse in (False, True):
        mods = [vnn.BilinearSamplerVolumetric(algo=a.algo << 8, vpt=a.vpt, hostAsync=True, reuseForwardInputs=reuse) for _ in range(2)]
        tickets = [None, None]

        def collect(k):
            if tickets[k] is not None:
                vnn.host_wait(tickets[k])
                tickets[k] = None
                return float(mods[k].output.view(-1)[0]) + float(mods[k].gradInput[1].view(-1)[0])

        def async_step(i):
            k = i & 1
            collect(k)                                                 # the step issued two steps ago: wait, read its result
            mods[k].updateOutput([vol_h, grid_h])
            mods[k].updateGradInput([vol_h, grid_h], go_h)
            tickets[k] = vnn.host_ticket()

        t = timed_loop(async_step, finish=lambda: (collect(0), collect(1), vnn.host_wait()))
        variants["async_reuse_inputs" if reuse else "async"] = {
            "ms_per_step": t * 1e3, "value": V * n_gpus / t,
            "h2d_bytes_per_step": (nv + ng + no) if reuse else 2 * (nv + ng) + no, "d2h_bytes_per_step": no + nv + ng,
            "note": "VOLSTN_HOST_ASYNC, two steps in flight: the next step's uploads overlap this step's downloads" + (" + input reuse" if reuse else "")}
        del mods
    h2d, d2h = 2 * (nv + ng) + no, no + nv + ng
    # the host link itself, every rank at the same time (pinned 256 MiB copies): one direction at a time, then both
    link = None
    try:
        nlk = 256 << 20
        hb, hb2 = torch.empty(nlk, dtype=torch.uint8).pin_memory(), torch.empty(nlk, dtype=torch.uint8).pin_memory()
        db, db2 = torch.empty(nlk, dtype=torch.uint8, device="cuda"), torch.empty(nlk, dtype=torch.uint8, device="cuda")
        s2 = torch.cuda.Stream()

        def rate(up, down):
            barrier()
            t0 = time.perf_counter()
            for _ in range(4):
                if up:
                    db.copy_(hb, non_blocking=True)
                if down:
                    with torch.cuda.stream(s2):
                        hb2.copy_(db2, non_blocking=True)
            torch.cuda.synchronize()
            return nlk * 4 / (time.perf_counter() - t0) / 1e9

        rate(True, True)
        up_gbs, down_gbs, both_gbs = rate(True, False), rate(False, True), rate(True, True)
        per_rank = None
        if distributed:
            tt = torch.tensor([up_gbs, down_gbs, both_gbs], dtype=torch.float64, device="cuda")
            allr = [torch.empty_like(tt) for _ in range(n_gpus)]
            dist.all_gather(allr, tt)
            per_rank = [[round(float(x), 2) for x in r] for r in allr]
        lb_ms = max(h2d / up_gbs, d2h / down_gbs) / 1e6
        # the step keeps both directions busy most of the time: every byte of it over this rank's link at the rate the link
        # carries when both directions run (all ranks measured at the same time) -- the ceiling the platform sets
        duplex_ms = (h2d + d2h) / (2 * both_gbs) / 1e6
        agg_both = 2 * (sum(r[2] for r in per_rank) if per_rank else both_gbs)
        link = {"h2d_gbs": up_gbs, "d2h_gbs": down_gbs, "each_way_gbs_when_both_run": both_gbs,
                "per_rank_h2d_d2h_both_gbs": per_rank, "aggregate_h2d_gbs": sum(r[0] for r in per_rank) if per_rank else up_gbs,
                "aggregate_d2h_gbs": sum(r[1] for r in per_rank) if per_rank else down_gbs,
                "aggregate_both_directions_gbs": agg_both,
                "step_gbs_all_ranks_both_directions": (h2d + d2h) * n_gpus / sec / 1e9,
                "frac_of_aggregate_duplex_ceiling": (h2d + d2h) * n_gpus / sec / 1e9 / agg_both,
                "duplex_bound_ms_per_step": duplex_ms,
                "lower_bound_ms_per_step": lb_ms, "frac_of_link_bound": lb_ms / (sec * 1e3),
                "numa_node_of_gpu": int(lib.volstn_b200_host_numa_node(local)),
                "note": "all ranks copy at the same time; the bound is max(upload bytes / h2d rate, download bytes / d2h rate) of this rank's link as measured with every other rank's link busy"}
        del hb, db, hb2, db2
    except Exception as e:  # noqa: BLE001
        link = {"error": repr(e)[:200]}
    return {"value": V * n_gpus / sec, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": sec * 1e3, "steps": e2e_steps, "link": link, "variants": variants,
            "path": "volstn_b200_host_sampler_forward/backward as lua/BilinearSamplerVolumetric.lua calls them: pinned host tensors -> sliced H2D / kernels / D2H on 3 streams -> host tensors; synchronous; both calls upload volume and grid (the reference re-reads its inputs)"}


def run_aux(a, torch, vnn, synth, _abi, peak):
    """SURVEY section 8 rows a7 / a8 and BASELINE.json config 2 (rank 0, N=1): the grid generator and Cumsum
    kernels and the 32^3 warp pipeline, each timed with an L2 flush before every timed launch."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn, reps=10, graph=True):
        """median device time of fn(); fn is captured into a CUDA graph first so that the Python / ctypes
        launch cost (tens of us, comparable to these kernels) is not inside the events"""
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        run = fn
        if graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            run = g.replay
            run()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    res = {}
    N, B = a.size, a.batch
    V = B * N ** 3
    gen = vnn.VolumetricGridGenerator(N, N, N)
    T = torch.eye(4, device="cuda").repeat(B, 1, 1) + 0.01 * torch.randn(B, 4, 4, device="cuda")
    gg = torch.randn((B, N, N, N, 4), device="cuda")
    ms = timed(lambda: gen.updateOutput(T))
    res["gridgen_forward"] = {"ms": ms, "achieved_gbs": 16 * V / (ms * 1e-3) / 1e9, "frac": 16 * V / (ms * 1e-3) / 1e9 / peak,
                              "algorithmic_bytes": 16 * V, "note": f"write-only 16 B/voxel, {N}^3 x batch {B}, graph replay, L2 flushed"}
    ms = timed(lambda: gen.updateGradInput(T, gg))
    res["gridgen_backward"] = {"ms": ms, "achieved_gbs": 16 * V / (ms * 1e-3) / 1e9, "frac": 16 * V / (ms * 1e-3) / 1e9 / peak,
                               "algorithmic_bytes": 16 * V, "note": "read-only 16 B/voxel, two-stage deterministic reduction"}
    del gg
    csz = 8
    cs = vnn.Cumsum()
    d = torch.rand((B, 3, csz, csz, csz), device="cuda")
    res["cumsum_forward_us"] = timed(lambda: cs.updateOutput(d)) * 1e3
    res["cumsum_backward_us"] = timed(lambda: cs.updateGradInput(d, d)) * 1e3
    res["cumsum_note"] = f"B x 3 x {csz}^3 = {d.numel() * 4 / 1024:.0f} KiB: launch-latency bound, reported in us of one graph-replayed launch (the reference needs ~15 launches)"

    # yardstick (NOT a parity oracle): ATen's 5-D grid_sample on the same problem -- same
    # mathematics, but NCDHW layout, (x, y, z) grid order, a different un-normalise expression and no 4th
    # grid component; forward + backward for both inputs.
    try:
        import torch.nn.functional as F
        rngy = np.random.default_rng(3)
        nb = min(B, 8)
        vol_y = torch.from_numpy(synth.volume("v2", nb, N, N, N, 1, rngy)).cuda().repeat(B // nb, 1, 1, 1, 1)[..., 0].unsqueeze(1).contiguous().requires_grad_(True)
        g_y = torch.from_numpy(synth.grid("g1", nb, N, N, N, rngy)[..., [2, 1, 0]].copy()).cuda().repeat(B // nb, 1, 1, 1, 1).contiguous().requires_grad_(True)
        go_y = torch.randn((B, 1, N, N, N), device="cuda")

        def aten_step():
            o = F.grid_sample(vol_y, g_y, mode="bilinear", padding_mode="zeros", align_corners=True)
            torch.autograd.grad(o, (vol_y, g_y), go_y)

        ms = timed(aten_step, reps=5, graph=False)
        res["aten_grid_sample_yardstick"] = {"ms": ms, "voxels_per_s": V / (ms * 1e-3),
                                             "note": "torch.nn.functional.grid_sample 5-D fwd+bwd, identity field, same 64^3 x batch; context only"}
        del vol_y, g_y, go_y
    except Exception as e:  # a yardstick must never break the bench
        res["aten_grid_sample_yardstick"] = {"error": repr(e)[:200]}

    # BASELINE.json config 2: Cumsum integration + grid generator + cage up-sample + source warp, 32^3, batch 32
    Np, Bp = 32, 32
    rng = np.random.default_rng(5)
    delta = 2.0 / (csz - 1)
    dd = torch.from_numpy(np.abs(delta * (1 + 0.25 * rng.standard_normal((Bp, 3, csz, csz, csz)))).astype(np.float32)).cuda()
    src = torch.from_numpy(synth.volume("v2", Bp, Np, Np, Np, 1, rng)).cuda()
    gop = torch.from_numpy(synth.grad_output(Bp, Np, Np, Np, 1, rng)).cuda()
    genp = vnn.VolumetricGridGenerator(Np, Np, Np)
    eye = torch.eye(4, device="cuda").repeat(Bp, 1, 1)
    csp, up, warp = vnn.Cumsum(), vnn.BilinearSamplerVolumetric(), vnn.BilinearSamplerVolumetric()
    cage_vol = torch.ones((Bp, csz, csz, csz, 4), device="cuda")
    gcage_in = torch.empty((Bp, 3, csz, csz, csz), device="cuda")

    def pipeline():
        cage = csp.updateOutput(dd)                                   # integrate the cage deltas
        cage_vol[..., :3] = (cage + (-1.0 - delta)).permute(0, 2, 3, 4, 1)
        field_i = genp.updateOutput(eye)                              # identity sampling field
        field = up.updateOutput([cage_vol, field_i])                  # cage up-sample -> B x 32^3 x 4 field
        warp.updateOutput([src, field])                               # source warp
        _, ggrid = warp.updateGradInput([src, field], gop)
        gcage, _ = up.updateGradInput([cage_vol, field_i], ggrid)
        gcage_in.copy_(gcage[..., :3].permute(0, 4, 1, 2, 3))
        csp.updateGradInput(dd, gcage_in)

    ms = timed(pipeline)
    Vp = Bp * Np ** 3
    res["pipeline_32"] = {"ms": ms, "voxels_per_s": Vp / (ms * 1e-3),
                          "workload": f"config 2: Cumsum + grid generator + cage up-sample (8^3 x 4 -> {Np}^3) + source warp, forward and backward, "
                                      f"{Np}^3 x batch {Bp}, CUDA graph replay, L2 flushed before each replay"}
    return res


def run_fused(a, torch, vnn, synth, _abi, peak):
    """SURVEY section 8 rows f1-f3 (rank 0, N=1): each fused / in-place module next to the chain of un-fused
    modules it replaces, same inputs, same results (tests/test_gpu_rowwarp.py), device-resident, K launches back to
    back between two CUDA events.  `bytes_per_voxel` are ALGORITHMIC bytes of the chain and of the fused call."""
    N, B, K = a.size, a.batch, max(5, min(a.steps, 20))
    V = B * N ** 3
    res = {"note": f"{N}^3 x batch {B}, C=1, fp32; ms per call; working sets exceed L2 (no flush) except pipeline_32 (flushed)"}
    rng = np.random.default_rng(11)
    nb = min(B, 8)

    def rep(x):
        return torch.from_numpy(x).cuda().repeat(B // nb, *([1] * (x.ndim - 1))).contiguous()

    def timed(fn, k=K):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    def timed_graph(fn, k=K):
        """one call captured in a CUDA graph and replayed k times: the device time of the call's launches without the
        Python / ctypes overhead of the module wrapper (~0.1 ms per call, which hides kernels shorter than that)"""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    vol = rep(synth.volume("v2", nb, N, N, N, 1, rng))
    go = torch.randn((B, N, N, N, 1), device="cuda")

    # ---- f1: cage up-sample + source warp ------------------------------------------------------------------
    csz = 8
    cage = torch.from_numpy(synth.alignet_cage(nb, csz, rng).astype(np.float32)).cuda().repeat(B // nb, 1, 1, 1, 1).contiguous()
    cage_vol = torch.ones((B, csz, csz, csz, 4), device="cuda")
    cage_vol[..., :3] = cage.permute(0, 2, 3, 4, 1)
    field_i = vnn.VolumetricGridGenerator(N, N, N).updateOutput(torch.eye(4, device="cuda").repeat(B, 1, 1))
    up, warp = vnn.BilinearSamplerVolumetric(onlyVolume=True), vnn.BilinearSamplerVolumetric(onlyGrid=True)
    fus = vnn.CageWarpVolumetric(onlyGrid=True)
    field = up.updateOutput([cage_vol, field_i])
    _, gfield = warp.updateGradInput([vol, field], go)
    c_f = timed(lambda: (up.updateOutput([cage_vol, field_i]), warp.updateOutput([vol, field])))
    c_b = timed(lambda: (warp.updateGradInput([vol, field], go), up.updateGradInput([cage_vol, field_i], gfield)), k=3)
    f_f = timed(lambda: fus.updateOutput([vol, cage]))
    f_b = timed(lambda: fus.updateGradInput([vol, cage], go))
    res["cage_warp"] = {
        "what": "models/alignet_2d.lua:136-147 lifted to 3-D: cage 8^3 x 3 -> field -> warp; backward to the cage only (source is data)",
        "chain": {"fwd_ms": c_f, "bwd_ms": c_b, "voxels_per_s": V / ((c_f + c_b) * 1e-3), "bytes_per_voxel": 56 + 72,
                  "calls": "sampler fwd (C=4 cage volume, constant identity field) + sampler fwd; sampler bwd onlyGrid + sampler bwd onlyVolume"},
        "fused": {"fwd_ms": f_f, "bwd_ms": f_b, "voxels_per_s": V / ((f_f + f_b) * 1e-3), "bytes_per_voxel": 8 + 8,
                  "calls": "volstn_b200_cage_warp_forward + _backward (ONLY_GRID)"},
        "speedup_fwd": c_f / f_f, "speedup_bwd": c_b / f_b, "speedup_step": (c_f + c_b) / (f_f + f_b),
        "hbm_not_allocated_mib": 2 * V * 16 / 2 ** 20}
    del field, field_i, gfield, cage_vol

    # ---- the reference's own graph: backward of the cage up-sample (a csz^3 x 4 volume under the full-resolution field) ------
    res["small_volume_backward"] = {
        "what": "BilinearSamplerVolumetric updateGradInput (gradInput[1] only) with a csz^3 x 4 cage volume and the identity field: "
                "direct kernel (VOLSTN_ALGO_DIRECT) vs the automatic choice (privatised, warp-aggregated)"}
    gf = torch.randn((B, N, N, N, 4), device="cuda")
    gf[..., 3] = 0
    fi = vnn.VolumetricGridGenerator(N, N, N).updateOutput(torch.eye(4, device="cuda").repeat(B, 1, 1))
    for c in (4, 8):
        cv = torch.rand((B, c, c, c, 4), device="cuda")
        md, ma = vnn.BilinearSamplerVolumetric(onlyVolume=True, algo=_abi.ALGO_DIRECT), vnn.BilinearSamplerVolumetric(onlyVolume=True)
        d_ms = timed(lambda: md.updateGradInput([cv, fi], gf), k=2)
        a_ms = timed(lambda: ma.updateGradInput([cv, fi], gf), k=5)
        res["small_volume_backward"][f"cage_{c}"] = {"direct_ms": d_ms, "auto_ms": a_ms, "speedup": d_ms / a_ms}
    del gf, fi

    # ---- f4: the whole training step of the warp head (forward + SmoothL1 + IoU + backward) in one kernel ----------
    target = rep(synth.volume("v2", nb, N, N, N, 1, rng))
    step = vnn.CageWarpSmoothL1Volumetric()

    def chain_step():   # the fused modules + the criterion as separate passes (torch's smooth_l1 as the stand-in for THNN)
        o = fus.updateOutput([vol, cage])
        x = o - target
        loss = torch.where(x.abs() < 1, 0.5 * x * x, x.abs() - 0.5).sum()
        g_o = x.clamp(-1, 1) / o.numel()
        fus.updateGradInput([vol, cage], g_o)
        return loss

    c_s = timed(chain_step)
    f_s = timed(lambda: step.forwardBackward([vol, cage], target))
    res["cage_warp_step"] = {
        "what": "model:forward + SmoothL1Criterion:forward/backward + computeIOU + model:backward (train.lua:123-126, model.lua:34, util.lua:110-120), backward to the cage",
        "fused_modules_plus_elementwise_criterion_ms": c_s, "one_kernel_ms": f_s, "speedup": c_s / f_s,
        "voxels_per_s": V / (f_s * 1e-3), "bytes_per_voxel": 8,
        "vs_unfused_module_chain": (res["cage_warp"]["chain"]["fwd_ms"] + res["cage_warp"]["chain"]["bwd_ms"] + (c_s - f_f - f_b)) / f_s}
    # VOLSTN_FAST_ARITH (tolerance mode: out within 1e-5 max|out| of the chain): the brick kernels of csrc/cagebrick.cu
    fast, fstep = vnn.CageWarpVolumetric(onlyGrid=True, fastArith=True), vnn.CageWarpSmoothL1Volumetric(fastArith=True)
    idc = rep(synth.alignet_cage(nb, 8, rng, jitter=0.0).astype(np.float32))
    ff_f, ff_b = timed(lambda: fast.updateOutput([vol, cage])), timed(lambda: fast.updateGradInput([vol, cage], go))
    ff_s, ff_si = timed(lambda: fstep.forwardBackward([vol, cage], target)), timed(lambda: fstep.forwardBackward([vol, idc], target))
    res["cage_warp"]["fast_arith"] = {"fwd_ms": ff_f, "bwd_ms": ff_b, "voxels_per_s": V / ((ff_f + ff_b) * 1e-3),
                                      "calls": "the same entry points with VOLSTN_FAST_ARITH (brick kernels)"}
    g_x = timed_graph(lambda: step.forwardBackward([vol, cage], target))
    g_s, g_si = timed_graph(lambda: fstep.forwardBackward([vol, cage], target)), timed_graph(lambda: fstep.forwardBackward([vol, idc], target))
    res["cage_warp_step"]["one_kernel_graph_ms"] = g_x
    res["cage_warp_step"]["fast_arith"] = {
        "one_kernel_ms": ff_s, "voxels_per_s": V / (ff_s * 1e-3), "speedup_over_exact": f_s / ff_s,
        "identity_cage_ms": ff_si, "identity_cage_voxels_per_s": V / (ff_si * 1e-3),
        "one_kernel_graph_ms": g_s, "identity_cage_graph_ms": g_si, "graph_speedup_over_exact": g_x / g_s,
        "graph_note": "the same calls captured in a CUDA graph and replayed: device time without the module wrapper's Python overhead",
        "note": "jittered cage (sheared field: ~18 L1 sectors per gather) and the identity cage ALIGNet starts from (weight 0 / bias del, models/alignet_2d.lua:118-119)"}
    del idc
    # the same one-kernel step for the plain sampler (a dense field tensor), identity field
    gridn0 = rep(synth.grid("g1", nb, N, N, N, rng))
    smp_og, sstep = vnn.BilinearSamplerVolumetric(onlyGrid=True), vnn.BilinearSamplerSmoothL1Volumetric()

    def chain_sampler_step():
        o = smp_og.updateOutput([vol, gridn0])
        x = o - target
        loss = torch.where(x.abs() < 1, 0.5 * x * x, x.abs() - 0.5).sum()
        smp_og.updateGradInput([vol, gridn0], x.clamp(-1, 1) / o.numel())
        return loss

    c_s = timed(chain_sampler_step)
    f_s = timed(lambda: sstep.forwardBackward([vol, gridn0], target))
    res["sampler_step"] = {
        "what": "BilinearSamplerVolumetric forward + SmoothL1Criterion forward/backward + computeIOU + backward (gradGrid only), grid tensor given",
        "modules_plus_elementwise_criterion_ms": c_s, "one_kernel_ms": f_s, "speedup": c_s / f_s,
        "voxels_per_s": V / (f_s * 1e-3), "bytes_per_voxel": 16 + 4 + 4 + 16,
        "frac_of_hbm_peak": 40 * V / (f_s * 1e-3) / 1e9 / peak}
    del target, gridn0

    # ---- f3: grid generator folded into the sampler ----------------------------------------------------------
    T = torch.eye(4, device="cuda").repeat(B, 1, 1) + 0.02 * torch.randn(B, 4, 4, device="cuda")
    gen, smp, afs = vnn.VolumetricGridGenerator(N, N, N), vnn.BilinearSamplerVolumetric(), vnn.AffineSamplerVolumetric(N, N, N)
    grid = gen.updateOutput(T)
    _, gg = smp.updateGradInput([vol, grid], go)
    c_f = timed(lambda: smp.updateOutput([vol, gen.updateOutput(T)]))
    c_b = timed(lambda: gen.updateGradInput(T, smp.updateGradInput([vol, grid], go)[1]))
    f_f = timed(lambda: afs.updateOutput([vol, T]))
    f_b = timed(lambda: afs.updateGradInput([vol, T], go))
    res["affine_sampler"] = {
        "what": "the projector of stn3dtorch/test.lua:15-25: VolumetricGridGenerator + BilinearSamplerVolumetric",
        "chain": {"fwd_ms": c_f, "bwd_ms": c_b, "voxels_per_s": V / ((c_f + c_b) * 1e-3), "bytes_per_voxel": 40 + 60},
        "fused": {"fwd_ms": f_f, "bwd_ms": f_b, "voxels_per_s": V / ((f_f + f_b) * 1e-3), "bytes_per_voxel": 8 + 12},
        "speedup_fwd": c_f / f_f, "speedup_bwd": c_b / f_b, "speedup_step": (c_f + c_b) / (f_f + f_b)}
    del grid, gg

    # the same through the HOST entry points (the loader-side augmentation runs on CPU tensors, dataset_2d.lua:285-311)
    if not a.no_e2e:
        Bh = min(B, 16)
        vol_h = vol[:Bh].cpu().pin_memory()
        T_h = T[:Bh].cpu().pin_memory()
        gen_h, smp_h, afs_h = vnn.VolumetricGridGenerator(N, N, N), vnn.BilinearSamplerVolumetric(reuseForwardInputs=False), vnn.AffineSamplerVolumetric(N, N, N)

        def wall(fn, k=3):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(k):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / k * 1e3

        ch = wall(lambda: smp_h.updateOutput([vol_h, gen_h.updateOutput(T_h)]))
        fu = wall(lambda: afs_h.updateOutput([vol_h, T_h]))
        Vh = Bh * N ** 3
        res["affine_sampler"]["host_forward"] = {
            "batch": Bh, "chain_ms": ch, "fused_ms": fu, "speedup": ch / fu,
            "chain_pcie_bytes": Vh * (16 + 4 + 16 + 4), "fused_pcie_bytes": Vh * 8 + Bh * 64,
            "path": "pinned host tensors in, host tensors out: host_gridgen_forward + host_sampler_forward vs host_affine_sampler_forward"}

    # ---- f2: channel-first tensors in place --------------------------------------------------------------------
    C = 4
    Bc = max(1, B // C)
    Vc = Bc * N ** 3
    volc = torch.randn((Bc, C, N, N, N), device="cuda")
    goc = torch.randn((Bc, C, N, N, N), device="cuda")
    gridc = rep(synth.grid("g2", nb, N, N, N, rng))[:Bc].permute(0, 4, 1, 2, 3).contiguous()
    cfm, clm = vnn.BilinearSamplerVolumetric(layout="BCDHW"), vnn.BilinearSamplerVolumetric()
    cl = lambda t: t.permute(0, 2, 3, 4, 1).contiguous()          # nn.Transpose = materialised copy (nn/Transpose.lua)
    cf = lambda t: t.permute(0, 4, 1, 2, 3).contiguous()
    c_f = timed(lambda: cf(clm.updateOutput([cl(volc), cl(gridc)])))
    c_b = timed(lambda: [cf(g) for g in clm.updateGradInput([cl(volc), cl(gridc)], cl(goc))], k=3)
    f_f = timed(lambda: cfm.updateOutput([volc, gridc]))
    f_b = timed(lambda: cfm.updateGradInput([volc, gridc], goc), k=3)
    res["channel_first"] = {
        "what": f"B x C x D x H x W tensors (C={C}, batch {Bc}, ALIGNet-like field g2): nn.Transpose copies + packed sampler vs the row kernels on permuted views",
        "chain": {"fwd_ms": c_f, "bwd_ms": c_b, "voxels_per_s": Vc / ((c_f + c_b) * 1e-3)},
        "in_place": {"fwd_ms": f_f, "bwd_ms": f_b, "voxels_per_s": Vc / ((f_f + f_b) * 1e-3),
                     "fwd_frac_of_hbm_peak": fwd_bytes(C) * Vc / (f_f * 1e-3) / 1e9 / peak,
                     "bwd_frac_of_hbm_peak": bwd_bytes(C) * Vc / (f_b * 1e-3) / 1e9 / peak},
        "speedup_fwd": c_f / f_f, "speedup_bwd": c_b / f_b, "speedup_step": (c_f + c_b) / (f_f + f_b)}
    del volc, goc, gridc

    # the ALIGNet case: single-channel volume (layouts coincide), channel-first FIELD as a conv net emits it -- the packed
    # kernels with planar component loads
    gridn = rep(synth.grid("g1", nb, N, N, N, rng))
    grid1 = gridn.permute(0, 4, 1, 2, 3).contiguous()
    vol1, go1 = vol.permute(0, 4, 1, 2, 3).contiguous(), go.permute(0, 4, 1, 2, 3).contiguous()
    n_f = timed(lambda: clm.updateOutput([vol, gridn]))
    n_b = timed(lambda: clm.updateGradInput([vol, gridn], go))
    c_f = timed(lambda: cf(clm.updateOutput([cl(vol1), cl(grid1)])))
    c_b = timed(lambda: [cf(g) for g in clm.updateGradInput([cl(vol1), cl(grid1)], cl(go1))], k=3)
    f_f = timed(lambda: cfm.updateOutput([vol1, grid1]))
    f_b = timed(lambda: cfm.updateGradInput([vol1, grid1], go1))
    res["channel_first_c1"] = {
        "what": f"B x 1 x D x H x W volume, B x 4 x D x H x W field (identity field g1), batch {B}: native-layout tensors / nn.Transpose copies + sampler / in place",
        "native_layout": {"fwd_ms": n_f, "bwd_ms": n_b},
        "chain": {"fwd_ms": c_f, "bwd_ms": c_b, "voxels_per_s": V / ((c_f + c_b) * 1e-3)},
        "in_place": {"fwd_ms": f_f, "bwd_ms": f_b, "voxels_per_s": V / ((f_f + f_b) * 1e-3),
                     "step_frac_of_hbm_peak": (fwd_bytes(1) + bwd_bytes(1)) * V / ((f_f + f_b) * 1e-3) / 1e9 / peak},
        "speedup_step": (c_f + c_b) / (f_f + f_b)}
    del gridn, grid1, vol1, go1

    # ---- BASELINE.json config 2 with the fused module: Cumsum -> +beta -> CageWarp, forward and backward ---------
    Np, Bp = 32, 32
    delta = 2.0 / (csz - 1)
    dd = torch.from_numpy(np.abs(delta * (1 + 0.25 * rng.standard_normal((Bp, 3, csz, csz, csz)))).astype(np.float32)).cuda()
    src = torch.from_numpy(synth.volume("v2", Bp, Np, Np, Np, 1, rng)).cuda()
    gop = torch.from_numpy(synth.grad_output(Bp, Np, Np, Np, 1, rng)).cuda()
    csp = vnn.Cumsum()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def pipeline_ms(cw):
        def pipeline():
            cage_p = csp.updateOutput(dd) + (-1.0 - delta)
            cw.updateOutput([src, cage_p])
            _, gcage = cw.updateGradInput([src, cage_p], gop)
            csp.updateGradInput(dd, gcage)

        for _ in range(3):
            pipeline()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            pipeline()
        g.replay()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    ms = pipeline_ms(vnn.CageWarpVolumetric(onlyGrid=True))
    ms_fast = pipeline_ms(vnn.CageWarpVolumetric(onlyGrid=True, fastArith=True))
    res["pipeline_32_fused"] = {"ms": ms, "voxels_per_s": Bp * Np ** 3 / (ms * 1e-3),
                                "fast_arith_ms": ms_fast, "fast_arith_voxels_per_s": Bp * Np ** 3 / (ms_fast * 1e-3),
                                "workload": f"config 2 with the fused module: Cumsum + beta + CageWarpVolumetric forward, backward to the cage, "
                                            f"Cumsum backward; {Np}^3 x batch {Bp}, CUDA graph replay, L2 flushed before each replay; "
                                            f"fast_arith = the same with VOLSTN_FAST_ARITH (brick kernels)"}
    return res


def _load_train_step():
    import importlib.util
    spec = importlib.util.spec_from_file_location("volstn_train_step", os.path.join(ROOT, "tools", "train_step.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_train_config(a, torch, dist, rank, n_gpus, distributed):
    """BASELINE.json config 3: ALIGNet-3D training step at 64^3, GLOBAL batch 64 sharded over the ranks (strong
    scaling), the one-kernel warp head, one NCCL all-reduce of the CNN gradients (tools/train_step.py)."""
    ts = _load_train_step()
    res = ts.run(steps=max(3, min(a.steps, 8)), warmup=3, global_batch=64, size=64, fused=2, rank=rank, world=n_gpus)
    torch.cuda.empty_cache()
    # the same step replayed as ONE CUDA graph (CNN, warp head, all-reduce, Adam): what is left when the ~150 launches of the
    # eager CNN are off the critical path -- it matters at 8 samples per GPU, where the eager step is launch-bound
    try:
        graphed = ts.run(steps=max(3, min(a.steps, 8)), warmup=3, global_batch=64, size=64, fused=2, rank=rank, world=n_gpus, graph=True)
        res["cuda_graph"] = {"ms_per_step": graphed["ms_per_step"], "samples_per_s": graphed["samples_per_s"], "loss": graphed["loss"]}
    except Exception as e:
        res["cuda_graph"] = {"error": repr(e)[:200]}
    torch.cuda.set_stream(torch.cuda.default_stream())
    torch.cuda.empty_cache()
    # the same step with the warp head in fast arithmetic (VOLSTN_FAST_ARITH: the brick kernels)
    try:
        fast = ts.run(steps=3, warmup=3, global_batch=64, size=64, fused=3, rank=rank, world=n_gpus)
        res["fast_arith_warp_head"] = {"ms_per_step": fast["ms_per_step"], "warp_ms": fast["stage_ms"]["warp"], "loss": fast["loss"]}
    except Exception as e:  # an extra measurement must never cost the contract line
        res["fast_arith_warp_head"] = {"error": repr(e)[:200]}
    torch.cuda.empty_cache()
    return res


def run_contention_config(a, torch, dist, rank, n_gpus, distributed):
    """BASELINE.json config 5: 128^3, C = 4, extreme / out-of-bounds field g4 (5 % exact corners, 1 % of all voxels on ONE
    point: gradVol contention), 8 samples per GPU; device-resident, CUDA events, max over ranks."""
    from alignet_b200 import nn as vnn, shard, synth
    N, C, B = 128, 4, 8
    rng = np.random.default_rng(50 + rank)
    vol = torch.from_numpy(synth.volume("v1", 2, N, N, N, C, rng)).cuda().repeat(B // 2, 1, 1, 1, 1).contiguous()
    grid = torch.from_numpy(synth.grid("g4", 2, N, N, N, rng)).cuda().repeat(B // 2, 1, 1, 1, 1).contiguous()
    go = torch.randn((B, N, N, N, C), device="cuda")
    m = vnn.BilinearSamplerVolumetric()
    for _ in range(3):
        m.updateOutput([vol, grid]); m.updateGradInput([vol, grid], go)
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    reps, f, b = 5, 0.0, 0.0
    for _ in range(reps):
        e[0].record(); m.updateOutput([vol, grid]); e[1].record(); m.updateGradInput([vol, grid], go); e[2].record()
        torch.cuda.synchronize()
        f += e[0].elapsed_time(e[1]) / reps; b += e[1].elapsed_time(e[2]) / reps
    if distributed:
        f, b = shard.max_over_ranks(f, device="cuda"), shard.max_over_ranks(b, device="cuda")
    peak, _ = measured_peak()
    V = B * N ** 3
    del vol, grid, go
    torch.cuda.empty_cache()
    return {"config": "BASELINE.json config 5: 128^3, C=4, field g4 (out of bounds + contention), 8 samples per GPU",
            "n_gpus": n_gpus, "fwd_ms": f, "bwd_ms": b, "voxels_per_s": V * n_gpus / ((f + b) * 1e-3), "scaling": "weak",
            "fwd_frac_of_hbm_peak": fwd_bytes(C) * V / (f * 1e-3) / 1e9 / peak, "bwd_frac_of_hbm_peak": bwd_bytes(C) * V / (b * 1e-3) / 1e9 / peak,
            "step_frac_of_hbm_peak": (fwd_bytes(C) + bwd_bytes(C)) * V / ((f + b) * 1e-3) / 1e9 / peak,
            "note": "working set 1.5 GiB per GPU (no L2 flush needed); parity of this configuration against the float and double oracles: tests/test_gpu_sampler.py::test_scatter_contention_against_double_oracle"}


def run_single_config(a):
    import torch
    import torch.distributed as dist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fn = run_train_config if a.config == "train" else run_contention_config
    res = fn(a, torch, dist, rank, world, world > 1)
    if rank == 0:
        emit(res)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_sweep(a, torch, vnn, synth, _abi):
    """BASELINE.json config 4: fwd+bwd over sizes x grids, GB/s against the roofline (rank 0, 1 GPU)."""
    peak, _ = measured_peak()
    rows = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for N, B, C in ((32, 32, 1), (32, 256, 1), (64, 64, 1), (64, 256, 1), (128, 8, 1), (128, 32, 1), (128, 8, 4)):
        for gk in ("g1", "g2", "g3", "g4"):
            rng = np.random.default_rng(N + B)
            vol = torch.from_numpy(synth.volume("v2", min(B, 8), N, N, N, C, rng)).cuda().repeat(B // min(B, 8), 1, 1, 1, 1)
            gs = synth.grid(gk, min(B, 8), N, N, N, rng)
            grid = torch.from_numpy(gs).cuda().repeat(B // min(B, 8), 1, 1, 1, 1)
            go = torch.randn((B, N, N, N, C), device="cuda")
            m = vnn.BilinearSamplerVolumetric(algo=a.algo << 8, vpt=a.vpt)
            for _ in range(3):
                m.updateOutput([vol, grid]); m.updateGradInput([vol, grid], go)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            f_ms, b_ms = [], []
            for _ in range(5):
                flush.zero_()
                e[0].record(); m.updateOutput([vol, grid]); e[1].record(); m.updateGradInput([vol, grid], go); e[2].record()
                torch.cuda.synchronize()
                f_ms.append(e[0].elapsed_time(e[1])); b_ms.append(e[1].elapsed_time(e[2]))
            V = B * N ** 3
            f, b = statistics.median(f_ms), statistics.median(b_ms)
            rows.append({"N": N, "B": B, "C": C, "grid": gk, "fwd_ms": f, "bwd_ms": b,
                         "voxels_per_s": V / ((f + b) * 1e-3),
                         "fwd_frac": fwd_bytes(C) * V / (f * 1e-3) / 1e9 / peak,
                         "bwd_frac": bwd_bytes(C) * V / (b * 1e-3) / 1e9 / peak,
                         "step_frac": (fwd_bytes(C) + bwd_bytes(C)) * V / ((f + b) * 1e-3) / 1e9 / peak})
            del vol, grid, go
    return rows


_JSON_OUT = None


def emit(line):
    """print the one JSON line on the REAL stdout (see main)"""
    print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)


def main():
    # stdout carries exactly one JSON line.  Libraries write there too (NCCL prints its version banner on
    # stdout from C whatever NCCL_DEBUG_FILE says), so file descriptor 1 is pointed at stderr for the whole
    # run and the JSON line goes to a duplicate of the original stdout.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--batch", type=int, default=64, help="samples per GPU")
    ap.add_argument("--channels", type=int, default=1)
    ap.add_argument("--grid", default="g1", choices=list(GRID_DESC))
    ap.add_argument("--algo", type=int, default=0, help="0 auto, 1 direct red, 2 smem-tile (VOLSTN_ALGO_*)")
    ap.add_argument("--vpt", type=int, default=0)
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-aux", action="store_true", help="skip the grid generator / Cumsum / 32^3 pipeline timings")
    ap.add_argument("--no-graph", action="store_true", help="launch the step from Python instead of replaying a CUDA graph")
    ap.add_argument("--no-fields", action="store_true", help="skip the extra per-field (g2, g3) device-resident timings")
    ap.add_argument("--no-extra", action="store_true", help="skip BASELINE.json configs 3 (training step) and 5 (contention)")
    ap.add_argument("--config", default="sampler", choices=["sampler", "train", "contention"],
                    help="sampler = the contract line; train / contention = only that configuration, as one JSON line")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    elif a.config != "sampler":
        run_single_config(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
