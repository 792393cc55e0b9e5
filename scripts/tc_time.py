"""Time the tensor-core integrator (maze 64k x 20 steps by default)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import closed_form as cf, synth
from tests.gpu_common import build_model, dev
task, n_steps, B = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else ("maze", 20, 65536)
mode = sys.argv[4] if len(sys.argv) > 4 else "fp16"
sd, ad = synth.TASK_DIMS[task]
spec = cf.NetSpec(sd, ad, 256, 4, 2)
m = build_model(spec, cf.make_params(spec, 5, 1.0))
s, a = synth.states_actions(6, B, sd, ad); probe = synth.rademacher(7, (B, ad))
st, at, pt = dev(s), dev(a), dev(probe)
flop = 2 * n_steps * (2 * ((sd + ad + 1) * 256 + 3 * 256 * 256 + 2 * 256 * ad) + 2 * (ad * 256 + 3 * 256 * 256 + 2 * 256 * ad))
f = lambda: m.score(st, at, n_steps=n_steps, probe=pt, precision=mode)
f(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"debug={os.environ.get('FRL_TC_DEBUG','0')} {task} B={B} n={n_steps} {mode}: {ms:.3f} ms {B/ms/1e3:.2f} M/s {B*flop/ms/1e9:.0f} TFLOP/s", flush=True)
