"""Quick GPU check of the tensor-core log-density kernel: error vs the fp32 kernel/oracle and timings."""
import sys, os, time
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

modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["fp16", "bf16"]
for task, gain, n_steps, B in (("hopper", 1.0, 32, 4096), ("hopper", 2.0, 32, 4096), ("maze", 2.0, 20, 4096), ("ant", 2.0, 32, 2048), ("sine", 2.0, 32, 1000)):
    sd, ad = synth.TASK_DIMS[task]
    spec = cf.NetSpec(sd, ad, 256, 4, 2)
    flat = cf.make_params(spec, 5, gain)
    m = build_model(spec, flat)
    s, a = synth.states_actions(6, B, sd, ad); probe = synth.rademacher(7, (B, ad))
    st, at, pt = dev(s), dev(a), dev(probe)
    le0, la0, _ = m.score(st, at, n_steps=n_steps, probe=pt, precision="fp32")
    for mode in modes:
        le, la, rw = m.score(st, at, n_steps=n_steps, probe=pt, precision=mode)
        torch.cuda.synchronize()
        for name, got, ref in (("E", le, le0), ("A", la, la0)):
            r = rel_err(got.cpu().numpy(), ref.cpu().numpy())
            print(f"{task} g={gain} {mode} {name}: med {np.median(r):.1e} p99 {np.quantile(r, .99):.1e} max {r.max():.1e} frac>2e-3 {(r > 2e-3).mean():.4f}", flush=True)
for task, n_steps, B in (("maze", 20, 65536), ("hopper", 32, 262144)):
    sd, ad = synth.TASK_DIMS[task]
    spec = cf.NetSpec(sd, ad, 256, 4, 2)
    m = build_model(spec, cf.make_params(spec, 5, 1.0))
    s, a = synth.states_actions(6, B, sd, ad); probe = synth.rademacher(7, (B, ad))
    st, at, pt = dev(s), dev(a), dev(probe)
    flop = 2 * n_steps * (2 * ((sd + ad + 1) * 256 + 3 * 256 * 256 + 2 * 256 * ad) + 2 * (ad * 256 + 3 * 256 * 256 + 2 * 256 * ad))
    for mode in ["fp32"] + modes:
        ms = timeit(lambda: m.score(st, at, n_steps=n_steps, probe=pt, precision=mode), n=3 if mode == "fp32" else 10)
        print(f"{task} B={B} n={n_steps} {mode}: {ms:.3f} ms  {B / ms / 1e3:.3f} M samples/s  {B * flop / ms / 1e9:.1f} TFLOP/s", flush=True)
