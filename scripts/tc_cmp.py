"""Compare a kernel variant (env FRL_TC_V2 / FRL_TC_V3 set by the caller) against the default tensor-core kernel and fp32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import closed_form as cf, synth
from tests.gpu_common import build_model, dev
task, n_steps, B = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else ("maze", 20, 1000)
sd, ad = synth.TASK_DIMS[task]
spec = cf.NetSpec(sd, ad, 256, 4, 2)
m = build_model(spec, cf.make_params(spec, 5, 1.0))
s, a = synth.states_actions(6, B, sd, ad); probe = synth.rademacher(7, (B, ad))
var = {k: os.environ.pop(k) for k in ("FRL_TC_V2", "FRL_TC_V3") if k in os.environ}
ref = m.score(dev(s), dev(a), n_steps=n_steps, probe=dev(probe), precision="fp16")
f32 = m.score(dev(s), dev(a), n_steps=n_steps, probe=dev(probe), precision="fp32")
os.environ.update(var)
got = m.score(dev(s), dev(a), n_steps=n_steps, probe=dev(probe), precision="fp16")
torch.cuda.synchronize()
for name, g, r, f in zip(("logp_E", "logp_A", "reward"), got, ref, f32):
    d1 = ((g - r).abs() / r.abs().clamp_min(1)).max().item()
    d2 = ((g - f).abs() / f.abs().clamp_min(1)).max().item()
    print(f"{name}: max rel vs default fp16 kernel {d1:.2e}, vs fp32 {d2:.2e}, finite {bool(torch.isfinite(g).all())}")
