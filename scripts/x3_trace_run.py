"""One traced launch of the fp16x3 integrator on the maze bench shape (FRL_TC_TRACE=file must be set); see
scripts/tc3x_trace_report.py for the stamps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import closed_form as cf, synth
from tests.gpu_common import build_model, dev
sd, ad = synth.TASK_DIMS["maze"]
spec = cf.NetSpec(sd, ad, 256, 4, 2)
m = build_model(spec, cf.make_params(spec, 5, 1.0))
B = 65536
s, a = synth.states_actions(6, B, sd, ad); probe = synth.rademacher(7, (B, ad))
m.score(dev(s), dev(a), n_steps=20, probe=dev(probe), precision="fp16x3")
torch.cuda.synchronize()
