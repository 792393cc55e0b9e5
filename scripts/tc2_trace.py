"""One traced launch of the N=256 integrator: FRL_TC_TRACE=<file> python scripts/tc2_trace.py [task n_steps B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import closed_form as cf, synth
from tests.gpu_common import build_model, dev
task, n_steps, B = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else ("maze", 20, 148 * 128)
sd, ad = synth.TASK_DIMS[task]
spec = cf.NetSpec(sd, ad, 256, 4, 2)
m = build_model(spec, cf.make_params(spec, 5, 1.0))
s, a = synth.states_actions(6, B, sd, ad); probe = synth.rademacher(7, (B, ad))
tr = os.environ.pop("FRL_TC_TRACE", None)
m.score(dev(s), dev(a), n_steps=n_steps, probe=dev(probe), precision="fp16"); torch.cuda.synchronize()
if tr: os.environ["FRL_TC_TRACE"] = tr
m.score(dev(s), dev(a), n_steps=n_steps, probe=dev(probe), precision="fp16"); torch.cuda.synchronize()
print("traced")
