"""Quick GPU check of the fp32-grade split-precision integrator (fp16x3): error vs the float64 oracle next to the
FP32 FFMA kernel's error on the same rows, ragged sizes, per-step probes / exact trace, and timings."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import closed_form as cf, synth
from tests.gpu_common import build_model, dev, rel_err

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def stat(tag, got, ref):
    r = rel_err(got, ref)
    print(f"{tag}: med {np.median(r):.1e} p99 {np.quantile(r, .99):.1e} max {r.max():.1e} frac>1e-4 {(r > 1e-4).mean():.4f}", flush=True)

what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "acc"):
    for task, gain, n_steps, B in (("maze", 2.0, 20, 1000), ("hopper", 1.0, 32, 2048), ("hopper", 2.0, 32, 2048), ("halfcheetah", 2.0, 32, 777),
                                   ("hand", 2.0, 32, 600), ("ant", 2.0, 32, 513), ("sine", 2.0, 32, 1000)):
        sd, ad = synth.TASK_DIMS[task]
        for heads in (2, 1):
            spec = cf.NetSpec(sd, ad, 256, 4, heads)
            flat = cf.make_params(spec, 5, gain)
            m = build_model(spec, flat)
            s, a = synth.states_actions(6, B, sd, ad); probe = synth.rademacher(7, (B, ad))
            st, at, pt = dev(s), dev(a), dev(probe)
            ref = cf.log_prob(spec, flat, s, a, "expert", probe, n_steps)
            for mode in ("fp32", "fp16x3"):
                if heads == 2:
                    got = m.log_prob(st, at, "expert", n_steps=n_steps, probe=pt, precision=mode)
                else:
                    got = m.log_prob(torch.cat([st, at], -1), n_steps=n_steps, probe=pt, precision=mode)
                torch.cuda.synchronize()
                stat(f"{task} heads={heads} g={gain} B={B} {mode}", got.cpu().numpy(), ref)
            if heads == 2 and task in ("hopper", "hand"):
                ps = synth.rademacher(8, (n_steps, B, ad))
                ref2 = cf.log_prob(spec, flat, s, a, "agent", ps, n_steps)
                got = m.log_prob(st, at, "agent", n_steps=n_steps, probe=dev(ps), precision="fp16x3")
                stat(f"{task} per-step probe agent fp16x3", got.cpu().numpy(), ref2)
                ref3 = cf.log_prob(spec, flat, s, a, "expert", None, n_steps, exact=True)
                got = m.log_prob(st, at, "expert", n_steps=n_steps, exact=True, precision="fp16x3")
                stat(f"{task} exact trace fp16x3", got.cpu().numpy(), ref3)
                le, la, rw = m.score(st, at, n_steps=n_steps, probe=pt, precision="fp16x3")
                one = m.log_prob(st, at, "expert", n_steps=n_steps, probe=pt, precision="fp16x3")
                print("  two-role launch == single-role launch:", bool(torch.equal(le, one)), " deterministic:",
                      bool(torch.equal(le, m.score(st, at, n_steps=n_steps, probe=pt, precision="fp16x3")[0])), flush=True)
if what in ("all", "time"):
    for task, n_steps, B in (("maze", 20, 65536), ("hopper", 32, 262144), ("halfcheetah", 32, 262144), ("ant", 32, 262144), ("hand", 32, 262144)):
        sd, ad = synth.TASK_DIMS[task]
        spec = cf.NetSpec(sd, ad, 256, 4, 2)
        m = build_model(spec, cf.make_params(spec, 5, 1.0))
        s, a = synth.states_actions(6, B, sd, ad); probe = synth.rademacher(7, (B, ad))
        st, at, pt = dev(s), dev(a), dev(probe)
        flop = 2 * n_steps * (2 * ((sd + ad + 1) * 256 + 3 * 256 * 256 + 2 * 256 * ad) + 2 * (ad * 256 + 3 * 256 * 256 + 2 * 256 * ad))
        for mode in (["fp32"] if task == "maze" else []) + ["fp16x3", "fp16"]:
            ms = timeit(lambda: m.score(st, at, n_steps=n_steps, probe=pt, precision=mode), n=3 if mode == "fp32" else 10)
            print(f"{task} B={B} n={n_steps} {mode}: {ms:.3f} ms  {B / ms / 1e3:.3f} M samples/s  {B * flop / ms / 1e9:.1f} alg TFLOP/s", flush=True)
