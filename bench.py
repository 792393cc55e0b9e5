#!/usr/bin/env python
"""Benchmark of the FLOWRIL density hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload maze] [--precision fp16]

A "step" is one pass of the hot path over one synthetic batch: log p_E, log p_A and the reward for B (s, a) rows
(two n_steps-step Euler integrations with the Hutchinson divergence; reference modril/flowril/flowril.py:175-208).
Default workload = BASELINE.json configs[1]: maze (s 6 + a 2), batch 65536, 20-step ODE, scrf net H=256 depth 4.
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (task dims key, batch per GPU, n_steps)
    "sine": ("sine", 65536, 32),
    "maze": ("maze", 65536, 20),
    "hopper": ("hopper", 262144, 32),
    "halfcheetah": ("halfcheetah", 262144, 32),
    "walker": ("walker", 262144, 32),
    "ant": ("ant", 1048576, 32),
    "ant111": ("ant111", 1048576, 32),
    "hand": ("hand", 1048576, 32),
    "pick": ("pick", 1048576, 32),
    "push": ("push", 1048576, 32),
}
HIDDEN, DEPTH = 256, 4


def algorithmic_flops_per_sample(s_dim, a_dim, n_steps, n_heads=2, H=HIDDEN, depth=DEPTH):
    """SURVEY 8(d) / BASELINE.md: GEMM multiply-adds of the reference formulation, 2 FLOP per MAC, no hoisting
    credit; Hutchinson with one probe: 2 roles * n_steps * (F_fwd + F_tan)."""
    d_in = a_dim + s_dim + 1
    f_fwd = 2 * (d_in * H + (depth - 1) * H * H + n_heads * H * a_dim)
    f_tan = 2 * (a_dim * H + (depth - 1) * H * H + n_heads * H * a_dim)
    return 2 * n_steps * (f_fwd + f_tan)


def measured_peak_tflops():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["bf16_tflops"]), "MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone, L2 flushed between steps)"
    except Exception:
        return 1590.0, "fallback 1.59 PFLOP/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def measured_hbm_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6550.0   # B200_PROFILING.md fallback


def train_hbm_bytes_per_row(d_in, a_dim, H=HIDDEN, depth=DEPTH, n_heads=2):
    """HBM bytes one training row costs in the tensor-core path (fp16 activation tiles, 1 mask bit per activation):
    forward: read the layer input + write its output (+ mask bits); loss: read h; dgrad: read dZ + bits, write dZ;
    wgrad: read dZ and the layer input of every layer."""
    k0 = -(-d_in // 16) * 16
    nop = -(-(n_heads * a_dim) // 16) * 16
    h, bits = 2 * H, H // 8
    fwd = (2 * k0 + h + bits) + (depth - 1) * (2 * h + bits) + (h + 2 * nop)
    dgrad = (2 * nop + bits + h) + (depth - 1) * (2 * h + bits)
    wgrad = (h + 2 * k0) + (depth - 1) * (2 * h) + (h + 2 * nop)
    return fwd + dgrad + wgrad


class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def make_inputs(task, B, n_steps, seed=0):
    import numpy as np

    from oracle import synth

    s_dim, a_dim = synth.TASK_DIMS[task]
    s, a = synth.states_actions(1000 + seed, B, s_dim, a_dim)
    probe = synth.rademacher(1234, (B, a_dim))  # same probe at every step and for both roles (scrf semantics)
    return s_dim, a_dim, s, a, probe


def make_trained_model(task, device, steps=200):
    """Default nn.Linear init under torch.manual_seed(0) + 200 CFM Adam steps (lr 1e-3) on the synthetic expert
    a = tanh(s[:, :a_dim]) + 0.1 N(0,1) so that the fields are non-trivial."""
    import numpy as np
    import torch

    from modril_b200.flowril import CoupledResidualFM, FusedAdam, train_step
    from oracle import synth

    s_dim, a_dim = synth.TASK_DIMS[task]
    torch.manual_seed(0)
    model = CoupledResidualFM(s_dim, a_dim, HIDDEN, depth=DEPTH).to(device)
    opt = FusedAdam([model.net], lr=1e-3)
    pool_s, pool_a = synth.states_actions(77, 8192, s_dim, a_dim)
    pool_e = synth.expert_actions(78, pool_s, a_dim)
    d_s, d_e, d_a = (torch.from_numpy(x).to(device) for x in (pool_s, pool_e, pool_a))
    g = torch.Generator(device="cpu").manual_seed(5)
    for it in range(steps):
        idx = torch.randint(0, 8192, (2, 256), generator=g).to(device)
        t = (torch.rand(2, 256, 1, generator=g) * (1 - 2e-3) + 1e-3).to(device)
        z = torch.randn(2, 256, a_dim, generator=g).to(device)
        train_step([dict(net=model.net, sign=1.0, s=d_s[idx[0]], a0=d_e[idx[0]], t=t[0], noise=z[0]),
                    dict(net=model.net, sign=-1.0, s=d_s[idx[1]], a0=d_a[idx[1]], t=t[1], noise=z[1])], opt, 1.0)
    torch.cuda.synchronize(device)
    return model


def cpu_port_rate(task, B_sample, n_steps, flat, threads, reps=1):
    """Scored samples/s of the torch-CPU port of the reference path (oracle/torch_port.py) on B_sample rows."""
    import torch

    from oracle import closed_form as cf
    from oracle import synth, torch_port

    torch.set_num_threads(threads)
    s_dim, a_dim = synth.TASK_DIMS[task]
    spec = cf.NetSpec(s_dim, a_dim, HIDDEN, DEPTH, 2)
    _, _, s, a, probe = make_inputs(task, B_sample, n_steps)
    st, at, pt = (torch.from_numpy(x) for x in (s, a, probe))
    ft = torch.from_numpy(flat).clone()
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        torch_port.score(spec, ft, ft, st, at, pt, n_steps)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return B_sample / best, best


def cpu_train_rate(task, B_sample, flat, threads, steps=3):
    """CFM train rows/s of the torch-CPU port (fm_loss x 2 -> backward -> clip_grad_norm_ -> Adam.step; flowril.py:311-330)
    on B_sample expert + B_sample agent rows per step."""
    import torch

    from oracle import closed_form as cf
    from oracle import synth, torch_port

    torch.set_num_threads(threads)
    s_dim, a_dim = synth.TASK_DIMS[task]
    spec = cf.NetSpec(s_dim, a_dim, HIDDEN, DEPTH, 2)
    s, a = synth.states_actions(3, B_sample, s_dim, a_dim)
    t, noise = synth.cfm_draws(4, B_sample, a_dim)
    st, at, tt, nt = (torch.from_numpy(x) for x in (s, a, t, noise))
    ft = torch.from_numpy(flat).clone().requires_grad_(True)
    opt = torch.optim.Adam([ft], lr=1e-4)
    batches = [(st, at, "expert", tt, nt), (st, at, "agent", tt, nt)]
    torch_port.train_step(spec, ft, opt, batches)          # warm-up
    t0 = time.perf_counter()
    for _ in range(steps):
        torch_port.train_step(spec, ft, opt, batches)
    dt = (time.perf_counter() - t0) / steps
    return 2 * B_sample / dt, dt


def cpu_baseline_leg(task, B, n_steps, flat):
    """`cpu_baseline` (scoring) and `train.cpu_baseline` of the GPU arm: the torch-CPU port of the reference op sequence on
    all host threads over a bounded sample (~12 s), plus the single-thread rate (the reference itself runs with
    torch.set_num_threads(1), rlf/run_settings.py:25-32) on a smaller one."""
    threads = os.cpu_count() or 1
    rate, _ = cpu_port_rate(task, 1024, n_steps, flat, threads, reps=2)
    Bs = int(min(4 * B, max(2048, rate * 12.0))) // 1024 * 1024  # ~12 s of CPU work
    rate, dt = cpu_port_rate(task, Bs, n_steps, flat, threads)
    rate1, dt1 = cpu_port_rate(task, 1024, n_steps, flat, 1)
    cpu = {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port",
           "sample": f"{Bs} of {B} rows, one pass ({dt:.1f} s), torch-CPU port of the reference op sequence "
                     f"(oracle/torch_port.py) with {threads} threads",
           "single_thread": {"value": rate1, "cores": 1, "sample": f"1024 rows, one pass ({dt1:.1f} s); the reference's own "
                                                                   f"setting (torch.set_num_threads(1))"}}
    trate, tdt = cpu_train_rate(task, 8192, flat, threads)
    cpu_train = {"value": trate, "unit": "samples/s", "cores": threads, "kind": "port",
                 "sample": f"3 updates of 2 x 8192 rows ({tdt:.2f} s each): fm_loss x 2, backward, clip_grad_norm_, Adam "
                           f"(oracle/torch_port.py::train_step) with {threads} threads"}
    return cpu, cpu_train


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (torch port, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import torch

    from oracle import closed_form as cf
    from oracle import synth

    task, B, n_steps = WORKLOADS[args.workload]
    s_dim, a_dim = synth.TASK_DIMS[task]
    spec = cf.NetSpec(s_dim, a_dim, HIDDEN, DEPTH, 2)
    flat = cf.make_params(spec, 0, 1.0)
    threads = os.cpu_count() or 1
    rate, _ = cpu_port_rate(task, 1024, n_steps, flat, threads, reps=2)  # calibration (first pass warms up)
    Bs = int(min(B, max(1024, rate * 4.0)))                               # ~4 s of CPU work per step
    Bs = Bs // 1024 * 1024
    times = []
    for i in range(args.warmup + args.steps):
        _, dt = cpu_port_rate(task, Bs, n_steps, flat, threads)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = Bs / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "scored_samples_per_sec", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # same workload / config keys as the GPU arm; the CPU steps a bounded sample of the batch
        "config": {"workload": f"{task}: s {s_dim} + a {a_dim}, {n_steps}-step ODE log-density (Hutchinson, caller "
                               f"probe), both roles + reward, scrf H={HIDDEN} depth={DEPTH}",
                   "batch_per_gpu": B, "n_steps": n_steps, "batch_per_step_sample": Bs,
                   "parallelism": f"host CPU, {threads} threads"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": f"{Bs} of {B} rows per step; torch-CPU port of the reference op sequence "
                                   f"(oracle/torch_port.py), torch {torch.__version__}"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from modril_b200.flowril import FusedAdam, score_pair_host, train_step
    from oracle import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    task, B, n_steps = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    s_dim, a_dim, s, a, probe = make_inputs(task, B, n_steps, seed=rank)
    model = make_trained_model(task, device)
    st, at, pt = (torch.from_numpy(x).to(device) for x in (s, a, probe))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)  # > 126 MB L2
    prec = args.precision

    def step():
        return model.score(st, at, n_steps=n_steps, probe=pt, precision=prec)

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- device-resident throughput (`value`) ----
    for _ in range(args.warmup):
        step()
        flush.zero_()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    # Keep the stream busy for ~50 ms first, so that the host (one Python process per GPU, sharing the box's cores at N = 8)
    # is ahead of the device for the whole region: an event interval then holds the step's kernels only, not the time the
    # device waited for the next launch (at N = 8 that idle time was 0.6 ms of a 1.8 ms step on the slowest rank).
    torch.cuda._sleep(100_000_000)
    torch.cuda.profiler.start()   # `ncu --profile-from-start off` then lists exactly the timed region's launches
    with ClockSampler(local) as clocks:
        for e0, e1 in ev:
            e0.record()
            out = step()
            e1.record()
            flush.zero_()
        barrier()
    torch.cuda.profiler.stop()
    total_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    tt = torch.tensor([total_ms], device=device, dtype=torch.float64)
    per_rank = None
    if world > 1:
        # every rank's own time and median SM clock under load (the aggregate is the MAX over ranks, as the contract says)
        mine = torch.tensor([total_ms / args.steps, float(clocks.summary()["sm_mhz"] or 0)], device=device, dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_per_step": [round(float(x[0]), 4) for x in allr], "sm_mhz": [int(x[1]) for x in allr]}
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
    ms_per_step = total_ms / args.steps
    value = world * B / (ms_per_step / 1e3)

    # ---- end to end through the host-buffer entry point (pinned host memory -> device -> pinned host) ----
    hs, ha, hp = (torch.from_numpy(x).pin_memory() for x in (s, a, probe))
    houts = tuple(torch.empty(B).pin_memory() for _ in range(3))

    def e2e_step():
        return score_pair_host(model.net, model.net, hs, ha, hp, n_steps=n_steps, precision=prec, out=houts)

    for _ in range(max(1, args.warmup)):
        e2e_step()
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        # the call returns after the results are in host memory, so its wall time is the end-to-end time of the step;
        # the L2 flush between steps (inputs arrive over PCIe anyway) stays outside the timed sum
        t0 = time.perf_counter()
        e2e_step()
        e2e_s += time.perf_counter() - t0
        flush.zero_()
        torch.cuda.synchronize(device)
    te = torch.tensor([e2e_s], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(te.item())
    e2e_check = float((houts[2] - out[2].cpu()).abs().max())

    # ---- parity spot check inside the bench: this mode vs the fp32 kernel on the first 4096 rows ----
    n_chk = min(4096, B)
    ref = model.score(st[:n_chk], at[:n_chk], n_steps=n_steps, probe=pt[:n_chk], precision="fp32")
    got = model.score(st[:n_chk], at[:n_chk], n_steps=n_steps, probe=pt[:n_chk], precision=prec)
    rel = ((got[0] - ref[0]).abs() / ref[0].abs().clamp_min(1.0)).cpu().numpy()
    parity = {"vs": "fp32 kernel, first %d rows, log p_E" % n_chk, "median_rel": float(np.median(rel)),
              "p99_rel": float(np.quantile(rel, 0.99)), "max_rel": float(rel.max())}

    # ---- other modes (kernel-only, 3 steps each) and the CFM train step, for context ----
    extra = {}
    for mode in ("fp32", "bf16", "fp16"):
        if mode == prec:
            continue
        model.score(st, at, n_steps=n_steps, probe=pt, precision=mode)
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 2 if mode == "fp32" else 5
        e0.record()
        for _ in range(reps):
            model.score(st, at, n_steps=n_steps, probe=pt, precision=mode)
        e1.record()
        torch.cuda.synchronize(device)
        extra[mode] = world * B / (e0.elapsed_time(e1) / reps / 1e3)

    Bt = 65536
    ts, ta = st[:Bt] if B >= Bt else st, at[:Bt] if B >= Bt else at
    Bt = ts.shape[0]
    g = torch.Generator(device="cpu").manual_seed(9 + rank)
    tt_ = (torch.rand(2, Bt, 1, generator=g) * (1 - 2e-3) + 1e-3).to(device)
    zz = torch.randn(2, Bt, a_dim, generator=g).to(device)
    grp = (dist.group.WORLD, world) if world > 1 else None
    # N = 1: one CUDA graph per update (all kernels of both roles + clip + Adam + weight re-pack).  N > 1: two graphs
    # (gradients | clip + Adam + re-pack) with the NCCL all-reduce launched between them, outside of any graph.
    # FRL_BENCH_NO_GRAPH=1 launches everything from Python instead.
    use_graph = os.environ.get("FRL_BENCH_NO_GRAPH", "0") != "1"
    opt = FusedAdam([model.net], lr=1e-4, capturable=use_graph)

    def tstep(prec_t, phase=None, losses=None):
        return train_step([dict(net=model.net, sign=1.0, s=ts, a0=ta, t=tt_[0], noise=zz[0]),
                           dict(net=model.net, sign=-1.0, s=ts, a0=ta, t=tt_[1], noise=zz[1])], opt, 1.0, world=grp,
                          precision=prec_t, phase=phase, losses=losses)

    train_rates = {}
    for prec_t in ("fp32", "fp16"):
        for _ in range(3):
            tstep(prec_t)
        run = lambda: tstep(prec_t)
        if use_graph and world == 1:
            torch.cuda.synchronize(device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                tstep(prec_t)
            run = graph.replay
            run()
        elif use_graph:
            torch.cuda.synchronize(device)
            g_grad, g_apply = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_grad):
                step_losses = tstep(prec_t, phase="grad")
            with torch.cuda.graph(g_apply):
                tstep(prec_t, phase="apply", losses=step_losses)

            def run(step_losses=step_losses, g_grad=g_grad, g_apply=g_apply):
                g_grad.replay()
                tstep(prec_t, phase="reduce", losses=step_losses)
                g_apply.replay()
            run()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_t = 10
        torch.cuda._sleep(40_000_000)   # ~20 ms: all n_t updates are enqueued before the first one runs (see the scoring loop)
        e0.record()
        for _ in range(n_t):
            run()
        e1.record()
        barrier()
        tms = torch.tensor([e0.elapsed_time(e1) / n_t], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        train_rates[prec_t] = world * 2 * Bt / (float(tms.item()) / 1e3)
    train_rate = train_rates["fp16"]

    # ---- BCF flow-matching policy (scope row f3), informational: acting latency and one behaviour-cloning update ----
    bcf_info = None
    if world == 1:
        from modril_b200.bcf import FlowPolicy, VectorNetwork, bc_step

        vnet = VectorNetwork(s_dim, a_dim).to(device)
        pol, bopt = FlowPolicy(vnet), FusedAdam([vnet], lr=1e-3, eps=1e-12)

        def timed(fn, n):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize(device)
            return e0.elapsed_time(e1) / n

        bcf_info = {"net": f"VectorNetwork s {s_dim} a {a_dim} H 256 T 128, FP32"}
        for Bb, n_it in ((32, 50), (128, 50), (65536, 5)):
            if Bb > B:
                continue
            bs_, ba_ = st[:Bb], at[:Bb]
            bx0, bt_ = torch.randn(Bb, a_dim, device=device), torch.rand(Bb, device=device)
            bcf_info[f"sample32_ms_B{Bb}"] = timed(lambda: pol.get_action(bs_, x0=bx0), n_it)
            bcf_info[f"bc_update_ms_B{Bb}"] = timed(lambda: bc_step(pol, bopt, bs_, ba_, x0=bx0, t=bt_), n_it)

    # ---- PPO consumer (scope row f4), informational: acting latency and one whole PPO.update of the reference's
    # default shape (128 steps x 32 processes, 4 epochs x 4 minibatches, --ppo-hidden-dim 64 x --ppo-layers 2)
    ppo_info = None
    if world == 1:
        from modril_b200.ppo import DistActorCritic, PPOUpdateGraph, ppo_update

        ppo_info = {}
        for tag, hid, rows in (("default_h64x2_rows4096", 64, 4096), ("h256x2_rows65536", 256, 65536)):
            if rows > B:
                continue
            pol_ac = DistActorCritic(s_dim, a_dim, hid, 2).to(device)
            popt = FusedAdam([pol_ac], lr=1e-3, eps=1e-12)
            po = st[:rows]
            out = pol_ac.get_action(po)
            pdata = (po, out["action"], out["action_log_probs"], out["value"], out["value"] + 0.1 * torch.randn_like(out["value"]),
                     torch.randn(rows, 1, device=device))
            ppo_info[f"act_ms_B32_{tag}"] = timed(lambda: pol_ac.get_action(po[:32]), 50)
            ppo_info[f"update_ms_{tag}"] = timed(lambda: ppo_update(pol_ac, popt, *pdata), 5)
            gopt = FusedAdam([pol_ac], lr=1e-3, eps=1e-12, capturable=True)
            graphed = PPOUpdateGraph(pol_ac, gopt, rows)
            ppo_info[f"update_graph_ms_{tag}"] = timed(lambda: graphed(*pdata), 5)
            ppo_info[f"update_rows_per_sec_{tag}"] = 4 * rows / (ppo_info[f"update_graph_ms_{tag}"] / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    d_in_t = a_dim + s_dim + 1
    f_fwd_train = 2 * (d_in_t * HIDDEN + (DEPTH - 1) * HIDDEN * HIDDEN + 2 * HIDDEN * a_dim)  # 3 F_fwd per row

    # ---- roofline of the dominant kernel (the fused integrator; the reward kernel is ~0.1 % of the step) ----
    flops = algorithmic_flops_per_sample(s_dim, a_dim, n_steps) * B
    peak, peak_src = measured_peak_tflops()
    achieved = flops / (ms_per_step / 1e3) / 1e12
    if prec == "fp32":
        peak, peak_src = 2 * 128 * 148 * 1.965e9 / 1e12, "FP32 FFMA peak 148 SM x 128 lanes x 2 x 1.965 GHz (nominal)"
    traffic, traffic_src = None, None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            ent = json.load(f).get(f"{args.workload}:{prec}:{B}")
        if ent:
            traffic, traffic_src = ent["bytes_per_launch"], ent["source"]
    except Exception:
        pass
    roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": B * ((s_dim + 2 * a_dim) * 4 + 12), "peak_source": peak_src,
            "kernel": "tc3::logprob_tc3_kernel (CTA pairs, tcgen05 cta_group::2)" if prec != "fp32" else "logprob_f32_kernel",
            "algorithmic_flops_per_sample": algorithmic_flops_per_sample(s_dim, a_dim, n_steps)}

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) ----
    cpu, cpu_train = None, None
    if world == 1 and not args.no_cpu:
        cpu, cpu_train = cpu_baseline_leg(task, B, n_steps, model.net.flat_params().detach().cpu().numpy().copy())

    line = {
        "metric": "scored_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": prec, "data": "synthetic",
        "config": {"workload": f"{task}: s {s_dim} + a {a_dim}, {n_steps}-step ODE log-density (Hutchinson, caller "
                               f"probe), both roles + reward, scrf H={HIDDEN} depth={DEPTH}",
                   "batch_per_gpu": B, "n_steps": n_steps, "parallelism": f"row-sharded x{world}, no collective",
                   "l2": "256 MiB buffer rewritten between timed steps (inputs are smaller than L2)",
                   "weights": "default init seed 0 + 200 CFM Adam steps on a synthetic expert"},
        "roofline": roof, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(world * B * (s_dim + 2 * a_dim) * 4),
                "d2h_bytes_per_step": int(world * B * 3 * 4), "api": "frl_score_host via score_pair_host, pinned host buffers",
                "max_abs_diff_vs_device_path": e2e_check},
        # per step: the fused integrator + the reward transform (+ the state-image pre-pass of wide-input nets)
        "gpu_launches": (3 if (prec != "fp32" and -(-(s_dim + a_dim + 1) // 16) * 16 > 32) else 2) * args.steps,
        "clocks": clocks.summary(), "per_rank": per_rank,
        "parity": parity,
        "other_modes_samples_per_sec": extra,
        "train": {"metric": "cfm_train_samples_per_sec", "value": train_rate, "unit": "samples/s",
                  "rows_per_step_per_gpu": 2 * Bt, "dtype": "fp16 operands / fp32 accumulate (tcgen05); fp32 clip + Adam",
                  "step": "frl_cfm_loss_grad_pair (both roles stacked: fwd + dgrad + wgrad) + [all-reduce] + frl_clip_adam + weight re-pack"
                          + (("; replayed as one CUDA graph" if world == 1 else
                              "; two CUDA graphs with the NCCL all-reduce between them") if use_graph else ""),
                  "fp32_ffma_samples_per_sec": train_rates["fp32"],
                  "algorithmic_tflops": train_rate * 3 * f_fwd_train / 1e12,
                  # what bounds it: the per-layer kernels stream fp16 activation tiles through HBM
                  "hbm": {"bytes_per_row": train_hbm_bytes_per_row(d_in_t, a_dim),
                          "achieved_gbs": train_rate * train_hbm_bytes_per_row(d_in_t, a_dim) / 1e9 / world,
                          "peak_gbs": measured_hbm_gbs(),
                          "frac": train_rate * train_hbm_bytes_per_row(d_in_t, a_dim) / 1e9 / world / measured_hbm_gbs()},
                  "frac_of_peak": train_rate * 3 * f_fwd_train / 1e12 / measured_peak_tflops()[0] / world,
                  "collective": "NCCL all-reduce of the flat grad" if world > 1 else "none",
                  "cpu_baseline": cpu_train},
        "bcf": bcf_info,
        "ppo": ppo_info,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="maze", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="fp16", choices=["fp32", "bf16", "fp16"])
    ap.add_argument("--batch", type=int, default=0, help="override rows per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
