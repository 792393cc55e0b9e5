#!/usr/bin/env python
"""Benchmark of the FLOWRIL density hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload maze] [--precision fp16x3]
                    [--metric score|train]

A "step" is one pass of the hot path over one synthetic batch: log p_E, log p_A and the reward for B (s, a) rows
(two n_steps-step Euler integrations with the Hutchinson divergence; reference modril/flowril/flowril.py:175-208).
Default workload = BASELINE.json configs[1]: maze (s 6 + a 2), batch 65536, 20-step ODE, scrf net H=256 depth 4, in the
default arithmetic of the module: fp16x3 = fp32-grade products on the tensor cores (the reference computes in fp32).  The
other BASELINE configs are timed in the `configs` block of the same line; the single-pass 16-bit modes in
`other_modes_samples_per_sec`.  `--metric train` makes the CFM training step (the part with a collective) the top-level
metric.  One JSON line is printed by rank 0.
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
CONFIGS_BLOCK = ("maze", "hopper", "halfcheetah", "ant", "hand")   # BASELINE.json configs[1..4]
WEIGHTS_DESC = "default init seed 0 + 200 CFM Adam steps on a synthetic expert"


def config_dict(task, s_dim, a_dim, n_steps, B, world):
    """`config` of the JSON line -- built by this one function for BOTH arms, so that they name the same configuration."""
    return {"workload": f"{task}: s {s_dim} + a {a_dim}, {n_steps}-step ODE log-density (Hutchinson, caller probe), both roles "
                        f"+ reward, scrf H={HIDDEN} depth={DEPTH}",
            "batch_per_gpu": B, "n_steps": n_steps, "parallelism": f"row-sharded x{world}, no collective",
            "l2": "256 MiB buffer rewritten between timed steps (inputs are smaller than L2)", "weights": WEIGHTS_DESC}


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


def training_draws(a_dim, steps=200):
    """Row indices, times and noise of the CFM warm-up updates (the same draws in the GPU arm and in the CPU arm)."""
    import torch

    g = torch.Generator(device="cpu").manual_seed(5)
    for _ in range(steps):
        idx = torch.randint(0, 8192, (2, 256), generator=g)
        t = torch.rand(2, 256, 1, generator=g) * (1 - 2e-3) + 1e-3
        z = torch.randn(2, 256, a_dim, generator=g)
        yield idx, t, z


def make_trained_model(task, device, steps=200):
    """Default nn.Linear init under torch.manual_seed(0) + 200 CFM Adam steps (lr 1e-3) on the synthetic expert
    a = tanh(s[:, :a_dim]) + 0.1 N(0,1) so that the fields are non-trivial."""
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
    for idx, t, z in training_draws(a_dim, steps):
        idx, t, z = idx.to(device), t.to(device), z.to(device)
        train_step([dict(net=model.net, sign=1.0, s=d_s[idx[0]], a0=d_e[idx[0]], t=t[0], noise=z[0]),
                    dict(net=model.net, sign=-1.0, s=d_s[idx[1]], a0=d_a[idx[1]], t=t[1], noise=z[1])], opt, 1.0)
    torch.cuda.synchronize(device)
    return model


# ---------------------------------------------------------------------------------------------------------------------
# The CPU arm.  With oracle/_ref (the reference's own modules, staged unmodified by oracle/build_ref.py) the timed code IS
# the reference: CoupledResidualFM.log_prob x 2 + the reward difference (flowril.py:183-191), fm_loss x 2 -> backward ->
# clip_grad_norm_ -> Adam.step (flowril.py:311-330).  Without it: the op-for-op torch port (oracle/torch_port.py).
# ---------------------------------------------------------------------------------------------------------------------
class CpuArm:
    def __init__(self, task, flat=None):
        import numpy as np
        import torch

        from oracle import build_ref
        from oracle import closed_form as cf
        from oracle import synth, torch_port

        self.task = task
        self.s_dim, self.a_dim = synth.TASK_DIMS[task]
        self.spec = cf.NetSpec(self.s_dim, self.a_dim, HIDDEN, DEPTH, 2)
        self.P = build_ref.load()
        self.kind = "reference" if self.P is not None else "port"
        self.model = None
        if self.P is not None:
            torch.manual_seed(0)
            self.model = self.P.CoupledResidualFM(self.s_dim, self.a_dim, HIDDEN, depth=DEPTH)
        if flat is None:   # the GPU arm's recipe on the CPU: default init (seed 0) + 200 CFM Adam steps, same draws
            if self.model is not None:
                flat = torch.cat([v.detach().reshape(-1) for v in self.model.state_dict().values()]).numpy().copy()
            else:
                flat = cf.make_params(self.spec, 0, 1.0)
            ft = torch.from_numpy(flat).clone().requires_grad_(True)
            opt = torch.optim.Adam([ft], lr=1e-3)
            pool_s, pool_a = synth.states_actions(77, 8192, self.s_dim, self.a_dim)
            pool_e = synth.expert_actions(78, pool_s, self.a_dim)
            d_s, d_e, d_a = (torch.from_numpy(x) for x in (pool_s, pool_e, pool_a))
            torch.set_num_threads(os.cpu_count() or 1)
            for idx, t, z in training_draws(self.a_dim):
                torch_port.train_step(self.spec, ft, opt, [(d_s[idx[0]], d_e[idx[0]], "expert", t[0], z[0]),
                                                           (d_s[idx[1]], d_a[idx[1]], "agent", t[1], z[1])])
            flat = ft.detach().numpy().copy()
        self.flat = np.asarray(flat, np.float32)
        if self.model is not None:
            sd = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in cf.unpack(self.spec, self.flat).items()}
            self.model.load_state_dict(sd)
        self.desc = ("the reference's own CoupledResidualFM (oracle/_ref, unmodified modril/flowril/{networks,pretrain}.py)"
                     if self.P is not None else "torch-CPU port of the reference op sequence (oracle/torch_port.py)")

    def score_rate(self, B_sample, n_steps, threads, reps=1):
        """Scored samples/s on B_sample rows: log p_E, log p_A, reward."""
        import torch

        from oracle import torch_port

        torch.set_num_threads(threads)
        _, _, s, a, probe = make_inputs(self.task, B_sample, n_steps)
        st, at, pt = (torch.from_numpy(x) for x in (s, a, probe))
        ft = torch.from_numpy(self.flat).clone()
        best = None
        for _ in range(reps):
            t0 = time.perf_counter()
            if self.model is not None:
                with torch.enable_grad():   # log_prob differentiates through the field for the divergence (pretrain.py:163-177)
                    le = self.model.log_prob(st, at, "expert", n_steps=n_steps)
                    la = self.model.log_prob(st, at, "agent", n_steps=n_steps)
                    (le - la).detach()
            else:
                torch_port.score(self.spec, ft, ft, st, at, pt, n_steps)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return B_sample / best, best

    def train_rate(self, B_sample, threads, steps=3):
        """CFM train rows/s: fm_loss x 2 -> backward -> clip_grad_norm_ -> Adam.step on B_sample expert + B_sample agent rows."""
        import torch

        from oracle import synth, torch_port

        torch.set_num_threads(threads)
        s, a = synth.states_actions(3, B_sample, self.s_dim, self.a_dim)
        st, at = torch.from_numpy(s), torch.from_numpy(a)
        if self.model is not None:
            import copy

            model = copy.deepcopy(self.model)
            opt = torch.optim.Adam(model.parameters(), lr=1e-4)

            def step():
                loss = model.fm_loss(st, at, "expert") + model.fm_loss(st, at, "agent")
                opt.zero_grad()
                loss.backward()
                torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
                opt.step()
        else:
            t, noise = synth.cfm_draws(4, B_sample, self.a_dim)
            tt, nt = torch.from_numpy(t), torch.from_numpy(noise)
            ft = torch.from_numpy(self.flat).clone().requires_grad_(True)
            opt = torch.optim.Adam([ft], lr=1e-4)
            batches = [(st, at, "expert", tt, nt), (st, at, "agent", tt, nt)]
            step = lambda: torch_port.train_step(self.spec, ft, opt, batches)
        step()          # warm-up
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = (time.perf_counter() - t0) / steps
        return 2 * B_sample / dt, dt


def cpu_baseline_leg(task, B, n_steps, flat):
    """`cpu_baseline` (scoring) and `train.cpu_baseline` of the GPU arm: the CPU arm on all host threads over a bounded
    sample (~12 s of CPU work), plus the single-thread rate (the reference itself runs with torch.set_num_threads(1),
    rlf/run_settings.py:25-32) on a smaller one.  The weights are the GPU arm's."""
    import torch

    threads = os.cpu_count() or 1
    arm = CpuArm(task, flat)
    rate, _ = arm.score_rate(1024, n_steps, threads, reps=2)
    Bs = max(2048, int(min(B, rate * 12.0)) // 1024 * 1024)   # ~12 s of CPU work, at most the batch
    rate, dt = arm.score_rate(Bs, n_steps, threads)
    rate1, dt1 = arm.score_rate(1024, n_steps, 1)
    cpu = {"value": rate, "unit": "samples/s", "cores": threads, "kind": arm.kind,
           "sample": f"the first {Bs} of the {B} rows, one pass ({dt:.1f} s); {arm.desc}; {threads} threads, torch {torch.__version__}",
           "single_thread": {"value": rate1, "cores": 1, "sample": f"1024 rows, one pass ({dt1:.1f} s); the reference's own "
                                                                   f"setting (torch.set_num_threads(1))"}}
    trate, tdt = arm.train_rate(8192, threads)
    cpu_train = {"value": trate, "unit": "samples/s", "cores": threads, "kind": arm.kind,
                 "sample": f"3 updates of 2 x 8192 rows ({tdt:.2f} s each): fm_loss x 2, backward, clip_grad_norm_, Adam; {arm.desc}; "
                           f"{threads} threads"}
    return cpu, cpu_train


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on all host threads (see CpuArm)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    task, B, n_steps = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    threads = os.cpu_count() or 1
    arm = CpuArm(task)           # same recipe for the weights as the GPU arm (default init seed 0 + 200 CFM Adam steps)
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if args.metric == "train":
        Bt = min(65536, B)
        Bs = 8192                # rows per role per step: a bounded sample of the 2 x Bt rows of the GPU arm's update
        times = []
        for i in range(args.warmup + args.steps):
            _, dt = arm.train_rate(Bs, threads, steps=1)
            if i >= args.warmup:
                times.append(dt)
        ms = 1e3 * sum(times) / len(times)
        value = 2 * Bs / (ms / 1e3)
        metric, sample = "cfm_train_samples_per_sec", f"{2 * Bs} of the {2 * Bt} rows of an update per step"
    else:
        rate, _ = arm.score_rate(1024, n_steps, threads, reps=2)          # calibration (first pass warms up)
        Bs = max(1024, int(min(B, rate * 4.0)) // 1024 * 1024)            # ~4 s of CPU work per step
        times = []
        for i in range(args.warmup + args.steps):
            _, dt = arm.score_rate(Bs, n_steps, threads)
            if i >= args.warmup:
                times.append(dt)
        ms = 1e3 * sum(times) / len(times)
        value = Bs / (ms / 1e3)
        metric, sample = "scored_samples_per_sec", f"the first {Bs} of the {B} rows per step"
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(task, arm.s_dim, arm.a_dim, n_steps, B, world),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": arm.kind,
                         "sample": f"{sample}; {arm.desc}; {threads} threads, torch {torch.__version__}"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


KERNELS = {"fp16x3": "tc3x::logprob_tc3x_kernel (CTA pairs, tcgen05 cta_group::2, 3-term fp16 split, main + small-terms accumulators)",
           "fp16": "tc3::logprob_tc3_kernel (CTA pairs, tcgen05 cta_group::2)", "bf16": "tc3::logprob_tc3_kernel (CTA pairs, tcgen05 cta_group::2)",
           "fp32": "logprob_f32_kernel (FFMA)"}


def measured_fp32_peak(device):
    """FP32 FFMA peak as MEASURED here: cuBLAS SGEMM 8192^3 with TF32 off, best of 5 (the denominator of the FFMA kernels)."""
    import torch

    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        a = torch.randn(8192, 8192, device=device)
        b = torch.randn(8192, 8192, device=device)
        torch.matmul(a, b)
        best = None
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize(device)
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        return 2 * 8192 ** 3 / (best / 1e3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def roofline_for(prec, achieved_tflops, fp32_peak=None):
    """Tensor roofline of a scoring mode.  fp16x3 evaluates every algorithmic product as THREE tensor-core products
    (x_hi W_hi + x_lo W_hi + x_hi W_lo), so its ceiling in algorithmic FLOP/s is a third of the 16-bit peak."""
    peak, src = measured_peak_tflops()
    if prec == "fp16x3":
        peak, src = peak / 3.0, "1/3 of " + src + ": three 16-bit MMA products per algorithmic product"
    elif prec == "fp32":
        peak, src = (fp32_peak, "cuBLAS SGEMM 8192^3 (TF32 off) measured in this run") if fp32_peak else \
                    (2 * 128 * 148 * 1.965e9 / 1e12, "FP32 FFMA nominal 148 SM x 128 lanes x 2 x 1.965 GHz")
    return {"bound": "tensor" if prec != "fp32" else "fp32 FFMA", "achieved": achieved_tflops, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved_tflops / peak, "peak_source": src, "kernel": KERNELS[prec]}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from modril_b200.flowril import FusedAdam, score_pair_host, train_step
    from oracle import closed_form as cf
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

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def max_over_ranks(x):
        tt = torch.tensor([x], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def time_scoring(m, xs, xa, xp, steps_, mode, n_iter, warm):
        """ms per scored batch: CUDA events around each call on the launching stream, L2 flushed between calls, MAX over ranks."""
        for _ in range(warm):
            m.score(xs, xa, n_steps=steps_, probe=xp, precision=mode)
            flush.zero_()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_iter)]
        barrier()
        torch.cuda._sleep(20_000_000)
        for e0, e1 in evs:
            e0.record()
            m.score(xs, xa, n_steps=steps_, probe=xp, precision=mode)
            e1.record()
            flush.zero_()
        barrier()
        return max_over_ranks(sum(e0.elapsed_time(e1) for e0, e1 in evs) / n_iter)

    def step():
        return model.score(st, at, n_steps=n_steps, probe=pt, precision=prec)

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
    per_rank = None
    if world > 1:
        # every rank's own time and median SM clock under load (the aggregate is the MAX over ranks, as the contract says)
        mine = torch.tensor([total_ms / args.steps, float(clocks.summary()["sm_mhz"] or 0)], device=device, dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_per_step": [round(float(x[0]), 4) for x in allr], "sm_mhz": [int(x[1]) for x in allr]}
    ms_per_step = max_over_ranks(total_ms) / args.steps
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
    e2e_value = world * B * args.steps / max_over_ranks(e2e_s)
    e2e_check = float((houts[2] - out[2].cpu()).abs().max())

    # ---- parity inside the bench: every mode against the float64 ORACLE on the first rows of the batch, bench weights ----
    parity = None
    if rank == 0:
        n_chk = min(2048, B)
        flat_np = model.net.flat_params().detach().cpu().numpy().astype(np.float64)
        spec = cf.NetSpec(s_dim, a_dim, HIDDEN, DEPTH, 2)
        ref_e = cf.log_prob(spec, flat_np, s[:n_chk], a[:n_chk], "expert", probe[:n_chk], n_steps)
        ref_a = cf.log_prob(spec, flat_np, s[:n_chk], a[:n_chk], "agent", probe[:n_chk], n_steps)
        parity = {"vs": f"float64 oracle (oracle/closed_form.py::log_prob), first {n_chk} rows, log p_E and log p_A, "
                        f"rel = |got - ref| / max(1, |ref|)", "modes": {}}
        for mode, bar in (("fp16x3", 1e-4), ("fp32", 1e-4), ("fp16", 2e-3), ("bf16", 2e-3)):
            got = model.score(st[:n_chk], at[:n_chk], n_steps=n_steps, probe=pt[:n_chk], precision=mode)
            rel = np.concatenate([np.abs(got[0].cpu().numpy() - ref_e) / np.maximum(1.0, np.abs(ref_e)),
                                  np.abs(got[1].cpu().numpy() - ref_a) / np.maximum(1.0, np.abs(ref_a))])
            parity["modes"][mode] = {"bar": bar, "median_rel": float(np.median(rel)), "p99_rel": float(np.quantile(rel, 0.99)),
                                     "max_rel": float(rel.max()), "rows_over_bar": int((rel > bar).sum()), "rows": int(rel.size)}
    barrier()

    # ---- other modes (kernel-only, a few steps each), for context ----
    extra = {}
    for mode in ("fp16x3", "fp16", "bf16", "fp32"):
        if mode == prec:
            continue
        ms_m = time_scoring(model, st, at, pt, n_steps, mode, 2 if mode == "fp32" else 5, 1)
        extra[mode] = world * B / (ms_m / 1e3)

    # ---- the other BASELINE.json configurations (rows per GPU, Euler steps), same measurement, in this mode and in fp16 ----
    configs = []
    if not args.no_configs:
        for ctask in CONFIGS_BLOCK:
            ck, cB, cn = WORKLOADS[ctask]
            if ctask == args.workload and cB == B:
                cm, (cs_, ca_, cp_) = model, (st, at, pt)
            else:
                _, _, xs, xa, xp = make_inputs(ck, cB, cn, seed=rank)
                cm = make_trained_model(ck, device)
                cs_, ca_, cp_ = (torch.from_numpy(x).to(device) for x in (xs, xa, xp))
            sd_, ad_ = synth.TASK_DIMS[ck]
            fl = algorithmic_flops_per_sample(sd_, ad_, cn)
            ent = {"workload": ctask, "dims": f"s {sd_} + a {ad_}", "batch_per_gpu": cB, "n_steps": cn, "modes": {}}
            for mode in dict.fromkeys((prec, "fp16")):
                ms_c = time_scoring(cm, cs_, ca_, cp_, cn, mode, 3, 1)
                rate = world * cB / (ms_c / 1e3)
                ent["modes"][mode] = {"ms_per_step": ms_c, "samples_per_sec": rate,
                                      "roofline_frac": roofline_for(mode, rate * fl / 1e12 / world)["frac"]}
            configs.append(ent)
            del cm, cs_, ca_, cp_
            torch.cuda.empty_cache()

    # ---- CFM training step (the part of the path with a collective) ----
    Bt = min(65536, B)
    ts, ta = st[:Bt], at[:Bt]
    g = torch.Generator(device="cpu").manual_seed(9 + rank)
    tt_ = (torch.rand(2, Bt, 1, generator=g) * (1 - 2e-3) + 1e-3).to(device)
    zz = torch.randn(2, Bt, a_dim, generator=g).to(device)
    # One CUDA graph per update (all kernels of both roles + [gradient all-reduce] + clip + Adam + weight re-pack).  N > 1:
    # the all-reduce is the device-side one-shot exchange over NVLink peer memory (modril_b200.dist.PeerAllReduce,
    # csrc/comm.cu: two kernels inside the graph).  FRL_BENCH_NCCL=1 takes the NCCL route instead: two graphs (gradients |
    # clip + Adam + re-pack) with ncclAllReduce launched between them from Python.  FRL_BENCH_NO_GRAPH=1 launches
    # everything from Python.
    use_graph = os.environ.get("FRL_BENCH_NO_GRAPH", "0") != "1"
    use_nccl = os.environ.get("FRL_BENCH_NCCL", "0") == "1"
    peer_comm = None
    peer_note = None
    if world > 1 and not use_nccl:
        from modril_b200.dist import PeerAllReduce

        try:
            peer_comm = PeerAllReduce(model.net.flat_params().numel() + 64, device=device)
        except RuntimeError as e:   # raised on every rank together: all of them take the NCCL route
            peer_comm, use_nccl, peer_note = None, True, str(e)[:300]
    grp = ((peer_comm if peer_comm is not None else dist.group.WORLD), world) if world > 1 else None
    opt = FusedAdam([model.net], lr=1e-4, capturable=use_graph)

    def tstep(prec_t, phase=None, losses=None, batch=None):
        xs, xa, xt, xz = batch if batch is not None else (ts, ta, tt_, zz)
        return train_step([dict(net=model.net, sign=1.0, s=xs, a0=xa, t=xt[0], noise=xz[0]),
                           dict(net=model.net, sign=-1.0, s=xs, a0=xa, t=xt[1], noise=xz[1])], opt, 1.0, world=grp,
                          precision=prec_t, phase=phase, losses=losses)

    train_rates, train_ms, train_run = {}, {}, {}
    for prec_t in ("fp32", "fp16"):
        for _ in range(3):
            tstep(prec_t)
        run = lambda: tstep(prec_t)
        if use_graph and (world == 1 or peer_comm is not None):
            torch.cuda.synchronize(device)
            barrier()
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

            def run(step_losses=step_losses, g_grad=g_grad, g_apply=g_apply, prec_t=prec_t):
                g_grad.replay()
                tstep(prec_t, phase="reduce", losses=step_losses)
                g_apply.replay()
            run()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_t = args.steps if (args.metric == "train" and prec_t == "fp16") else 10
        torch.cuda._sleep(40_000_000)   # ~20 ms: all n_t updates are enqueued before the first one runs (see the scoring loop)
        e0.record()
        for _ in range(n_t):
            run()
        e1.record()
        barrier()
        train_ms[prec_t] = max_over_ranks(e0.elapsed_time(e1) / n_t)
        train_rates[prec_t] = world * 2 * Bt / (train_ms[prec_t] / 1e3)
        train_run[prec_t] = run
    train_rate = train_rates["fp16"]
    # end to end: the update's inputs come from pinned host memory every step and the two losses go back to the host.
    # Public API only: two TrainStepGraphs over two sets of static device inputs; the copy of batch i + 1 runs on a side
    # stream while update i computes, and the loss read-back of update i is the step's result.
    h_in = [x.cpu().pin_memory() for x in (ts, ta, tt_, zz)]
    train_e2e_s, e2e_api = 0.0, None
    if use_graph and (world == 1 or peer_comm is not None):
        from modril_b200.flowril import TrainStepGraph

        copy_stream = torch.cuda.Stream(device)
        sets = []
        for k in range(2):
            d = [torch.empty_like(x, device=device) for x in h_in]
            for dst, src in zip(d, h_in):
                dst.copy_(src)
            terms_k = [dict(net=model.net, sign=1.0, s=d[0], a0=d[1], t=d[2][0], noise=d[3][0]),
                       dict(net=model.net, sign=-1.0, s=d[0], a0=d[1], t=d[2][1], noise=d[3][1])]
            barrier()
            sets.append((d, TrainStepGraph(terms_k, opt, 1.0, world=grp, precision="fp16"), torch.cuda.Event(), torch.cuda.Event()))
        torch.cuda.synchronize(device)
        for _, _, ev_copied, ev_consumed in sets:
            ev_copied.record()
            ev_consumed.record()
        main = torch.cuda.current_stream(device)
        for i in range(3 + 10):
            if i == 3:
                torch.cuda.synchronize(device)
                barrier()
                t0 = time.perf_counter()
            d, graph_k, ev_copied, ev_consumed = sets[i % 2]
            main.wait_event(ev_copied)
            losses_ = graph_k()
            ev_consumed.record(main)
            nd, _, nev_copied, nev_consumed = sets[(i + 1) % 2]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(nev_consumed)
                for dst, src in zip(nd, h_in):
                    dst.copy_(src, non_blocking=True)
                nev_copied.record(copy_stream)
            [float(x) for x in losses_]
        train_e2e_s = (time.perf_counter() - t0) / 2.0   # (10 timed updates; the expression below assumes 5)
        e2e_api = ("TrainStepGraph x 2 (double-buffered static inputs): H2D of batch i + 1 from pinned memory on a copy stream "
                   "while update i runs; the two losses of every update read back")
    else:
        d_in = [torch.empty_like(x, device=device) for x in h_in]
        for i in range(3 + 5):
            t0 = time.perf_counter()
            for dst, src in zip(d_in, h_in):
                dst.copy_(src, non_blocking=True)
            losses_ = tstep("fp16", batch=tuple(d_in))
            [float(x) for x in losses_]
            if i >= 3:
                train_e2e_s += time.perf_counter() - t0
        e2e_api = "train_step on device copies refreshed from pinned host memory every step; the two losses read back"
    train_e2e = world * 2 * Bt * 5 / max_over_ranks(train_e2e_s)
    train_h2d = int(world * sum(x.numel() * 4 for x in h_in))

    # ---- BCF flow-matching policy (scope row f3), informational: acting latency and one behaviour-cloning update ----
    bcf_info = None
    if world == 1 and not args.no_extras:
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
    if world == 1 and not args.no_extras:
        from modril_b200.ppo import DistActorCritic, PPOUpdateGraph, ppo_update

        ppo_info = {}
        for tag, hid, rows in (("default_h64x2_rows4096", 64, 4096), ("h256x2_rows65536", 256, 65536)):
            if rows > B:
                continue
            pol_ac = DistActorCritic(s_dim, a_dim, hid, 2).to(device)
            popt = FusedAdam([pol_ac], lr=1e-3, eps=1e-12)
            po = st[:rows]
            out_ac = pol_ac.get_action(po)
            pdata = (po, out_ac["action"], out_ac["action_log_probs"], out_ac["value"],
                     out_ac["value"] + 0.1 * torch.randn_like(out_ac["value"]), torch.randn(rows, 1, device=device))
            ppo_info[f"act_ms_B32_{tag}"] = timed(lambda: pol_ac.get_action(po[:32]), 50)
            ppo_info[f"update_ms_{tag}"] = timed(lambda: ppo_update(pol_ac, popt, *pdata), 5)
            gopt = FusedAdam([pol_ac], lr=1e-3, eps=1e-12, capturable=True)
            graphed = PPOUpdateGraph(pol_ac, gopt, rows)
            ppo_info[f"update_graph_ms_{tag}"] = timed(lambda: graphed(*pdata), 5)
            ppo_info[f"update_rows_per_sec_{tag}"] = 4 * rows / (ppo_info[f"update_graph_ms_{tag}"] / 1e3)

    fp32_peak = measured_fp32_peak(device) if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    d_in_t = a_dim + s_dim + 1
    f_fwd_train = 2 * (d_in_t * HIDDEN + (DEPTH - 1) * HIDDEN * HIDDEN + 2 * HIDDEN * a_dim)  # 3 F_fwd per row

    # ---- roofline of the dominant kernel (the fused integrator; the reward kernel is ~0.1 % of the step) ----
    fl_sample = algorithmic_flops_per_sample(s_dim, a_dim, n_steps)
    roof = roofline_for(prec, fl_sample * B / (ms_per_step / 1e3) / 1e12, fp32_peak)
    traffic_tab = {}
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic_tab = json.load(f)
    except Exception:
        pass
    ent = traffic_tab.get(f"{args.workload}:{prec}:{B}")
    roof.update({"traffic": ent["bytes_per_launch"] if ent else None, "traffic_source": ent["source"] if ent else None,
                 "algorithmic_bytes_per_launch": B * ((s_dim + 2 * a_dim) * 4 + 12), "algorithmic_flops_per_sample": fl_sample})

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) ----
    cpu, cpu_train = None, None
    if world == 1 and not args.no_cpu:
        cpu, cpu_train = cpu_baseline_leg(task, B, n_steps, model.net.flat_params().detach().cpu().numpy().copy())

    # HBM traffic of one update: MEASURED dram bytes (ncu launch list of scripts/train_time.py, profiles/traffic.json) when a
    # capture of this shape is committed; the modelled figure of the per-layer design is kept beside it for reference
    tent = traffic_tab.get(f"train:{args.workload}:fp16:{2 * Bt}")
    train_bytes_row = tent["bytes_per_launch"] / (2 * Bt) if tent else None
    train_block = {
        "metric": "cfm_train_samples_per_sec", "value": train_rate, "unit": "samples/s", "ms_per_step": train_ms["fp16"],
        "rows_per_step_per_gpu": 2 * Bt, "dtype": "fp16 operands / fp32 accumulate (tcgen05); fp32 clip + Adam",
        "step": "frl_cfm_loss_grad_pair (both roles stacked) + [all-reduce] + frl_clip_adam + weight re-pack"
                + (("; replayed as one CUDA graph" if (world == 1 or peer_comm is not None) else
                    "; two CUDA graphs with the NCCL all-reduce between them") if use_graph else ""),
        "fp32_ffma_samples_per_sec": train_rates["fp32"],
        "algorithmic_tflops": train_rate * 3 * f_fwd_train / 1e12,
        "roofline": {"bound": "tensor", "achieved": train_rate * 3 * f_fwd_train / 1e12 / world, "peak": measured_peak_tflops()[0],
                     "unit": "TFLOP/s", "frac": train_rate * 3 * f_fwd_train / 1e12 / measured_peak_tflops()[0] / world,
                     "algorithmic_flops_per_row": 3 * f_fwd_train},
        "hbm": {"bytes_per_row_measured": train_bytes_row, "source": tent["source"] if tent else None,
                "achieved_gbs": train_rate * train_bytes_row / 1e9 / world if tent else None, "peak_gbs": measured_hbm_gbs(),
                "frac": train_rate * train_bytes_row / 1e9 / world / measured_hbm_gbs() if tent else None,
                "algorithmic_bytes_per_row": (s_dim + 2 * a_dim + 1) * 4},
        "e2e": {"value": train_e2e, "unit": "samples/s", "h2d_bytes_per_step": train_h2d, "d2h_bytes_per_step": 8 * world,
                "api": e2e_api},
        "collective": "none" if world == 1 else (
            "device-side LL two-shot all-reduce of the flat fp32 gradient over NVLink peer memory (csrc/comm.cu), one kernel inside the update's graph"
            if peer_comm is not None else "NCCL all-reduce of the flat grad, launched from Python between two graphs"),
        "collective_timeouts": peer_comm.timeouts() if peer_comm is not None else None,
        "collective_fallback": peer_note,
        "cpu_baseline": cpu_train}

    line = {
        "metric": "scored_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": prec, "data": "synthetic",
        "config": config_dict(task, s_dim, a_dim, n_steps, B, world),
        "roofline": roof, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(world * B * (s_dim + 2 * a_dim) * 4),
                "d2h_bytes_per_step": int(world * B * 3 * 4), "api": "frl_score_host via score_pair_host, pinned host buffers",
                "max_abs_diff_vs_device_path": e2e_check},
        # per step: the fused integrator + the reward transform (+ the state-image pre-pass of wide-input nets)
        "gpu_launches": (3 if (prec != "fp32" and -(-(s_dim + a_dim + 1) // 16) * 16 > 32) else 2) * args.steps,
        "clocks": clocks.summary(), "per_rank": per_rank,
        "dtype_note": {"fp16x3": "3-term fp16 operand split on tcgen05, fp32 accumulate / state / trace: fp32-grade, the module default",
                       "fp16": "fp16 operands, one pass", "bf16": "bf16 operands, one pass", "fp32": "FP32 FFMA"}[prec],
        "parity": parity,
        "other_modes_samples_per_sec": extra,
        "configs": configs,
        "train": train_block,
        "bcf": bcf_info,
        "ppo": ppo_info,
    }
    if args.metric == "train":
        # the communicating part of the path as the top-level metric (for the 1 -> 8 scaling run): same line, keys swapped
        line.update({"metric": "cfm_train_samples_per_sec", "value": train_rate, "ms_per_step": train_ms["fp16"],
                     "dtype": "fp16", "roofline": dict(train_block["roofline"], traffic=tent["bytes_per_launch"] if tent else None),
                     "e2e": train_block["e2e"], "cpu_baseline": cpu_train, "scored": {"value": value, "ms_per_step": ms_per_step},
                     "config": dict(config_dict(task, s_dim, a_dim, n_steps, B, world),
                                    workload=f"{task}: s {s_dim} + a {a_dim}, CFM update of 2 x {Bt} rows per GPU (expert + agent), "
                                             f"clip 1.0 + Adam, scrf H={HIDDEN} depth={DEPTH}",
                                    parallelism=f"data parallel x{world}" + (", all-reduce of the flat gradient" if world > 1 else "")),
                     # our kernels per update: weight images, X0 tiles, 5 forward + heads/loss + 4 dgrad GEMM launches, wgrad,
                     # reduction, [peer all-reduce], sum of squares, clip + Adam
                     "gpu_launches": args.steps * (17 + (1 if peer_comm is not None else 0))})
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
    ap.add_argument("--precision", default="fp16x3", choices=["fp16x3", "fp32", "bf16", "fp16"])
    ap.add_argument("--metric", default="score", choices=["score", "train"],
                    help="top-level metric: scored samples/s (default) or CFM train samples/s (the path with a collective)")
    ap.add_argument("--no-configs", action="store_true", help="skip the block with the other BASELINE configurations")
    ap.add_argument("--no-extras", action="store_true", help="skip the informational BCF / PPO blocks")
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
