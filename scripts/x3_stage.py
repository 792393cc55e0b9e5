import sys, os
sys.path.insert(0, "/root/repo")
import torch
from oracle import closed_form as cf, synth
from tests.gpu_common import build_model, dev
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
sd, ad = synth.TASK_DIMS["maze"]
spec = cf.NetSpec(sd, ad, 256, 4, 2)
m = build_model(spec, cf.make_params(spec, 5, 1.0))
B, n_steps = 65536, 20
s, a = synth.states_actions(6, B, sd, ad); probe = synth.rademacher(7, (B, ad))
st, at, pt = dev(s), dev(a), dev(probe)
for mode in ("fp16x3", "fp16"):
    ms = timeit(lambda: m.score(st, at, n_steps=n_steps, probe=pt, precision=mode))
    print(os.environ.get("FRL_TC3_STAGES"), mode, f"{ms:.3f} ms", flush=True)
