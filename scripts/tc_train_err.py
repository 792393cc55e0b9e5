"""Error of the tensor-core (fp16) CFM gradient vs the float64 oracle for default-scale and stress nets (GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import closed_form as cf, synth
from tests.gpu_common import build_model, dev

def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum())

for dims in [(11, 3, 256, 4, 2), (17, 6, 256, 4, 2), (1, 1, 128, 4, 2), (6, 2, 128, 3, 1), (132, 8, 256, 4, 2)]:
    for gain in (1.0, 2.0):
        for B in (200, 4096):
            spec = cf.NetSpec(*dims)
            flat = cf.make_params(spec, 3, gain)
            s, a = synth.states_actions(8, B, spec.s_dim, spec.a_dim)
            t, noise = synth.cfm_draws(9, B, spec.a_dim)
            ref_loss, ref_grad = cf.fm_loss_and_grad(spec, flat, s, a, "expert", t, noise)
            out = {}
            for prec in ("fp32", "fp16"):
                m = build_model(spec, flat)
                kw = dict(t=dev(t), noise=dev(noise), precision=prec)
                loss = m.fm_loss(dev(s), dev(a), "expert", **kw) if spec.n_heads == 2 else \
                    m.fm_loss(dev(s), dev(a), dequant=torch.zeros(B, spec.a_dim, device="cuda"), **kw)
                loss.backward()
                g = torch.cat([p.grad.reshape(-1) for p, _, _ in m.net.param_slices()]).cpu().numpy()
                out[prec] = (loss.item(), g)
            print(dims, "gain", gain, "B", B, "loss rel err fp32 %.1e fp16 %.1e | grad relL2 fp32 %.1e fp16 %.1e" % (
                abs(out["fp32"][0] - ref_loss) / abs(ref_loss), abs(out["fp16"][0] - ref_loss) / abs(ref_loss),
                rel(out["fp32"][1], ref_grad), rel(out["fp16"][1], ref_grad)), flush=True)
