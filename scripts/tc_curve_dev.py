"""Deviation of the tensor-core (fp16-operand) training curve from the reference's 1000-step goldens, for the tolerance
statement of tests/test_gpu_train_tc.py::test_tc_training_curve_vs_reference.  usage: tc_curve_dev.py [fp16|fp32]"""
import glob
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from modril_b200.flowril import FusedAdam, train_step  # noqa: E402
from oracle import closed_form as cf  # noqa: E402
from oracle import synth  # noqa: E402
from tests.gpu_common import build_model, dev, spec_of  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "train_0[04]_*1000steps.npz"))):
    g = np.load(path)
    spec = spec_of(g["meta"])
    steps, B, seed = (int(x) for x in g["meta"][5:8])
    m = build_model(spec, cf.make_params(spec, seed, 1.0))
    opt = FusedAdam([m.net], lr=float(g["lr"]))
    pool_s, _ = synth.states_actions(seed + 1, 4096, spec.s_dim, spec.a_dim)
    pool_aE = synth.expert_actions(seed + 2, pool_s, spec.a_dim)
    pool_aA = np.clip(synth.rng(seed + 3).standard_normal((4096, spec.a_dim)), -1, 1).astype(np.float32)
    d_s, d_aE, d_aA = dev(pool_s), dev(pool_aE), dev(pool_aA)
    log = []
    for it in range(steps):
        r = synth.rng(seed + 100 + it)
        iE, iA = dev(r.integers(0, 4096, B)), dev(r.integers(0, 4096, B))
        tE, nE = synth.cfm_draws(seed + 10000 + 2 * it, B, spec.a_dim)
        tA, nA = synth.cfm_draws(seed + 10001 + 2 * it, B, spec.a_dim)
        terms = [dict(net=m.net, sign=1.0, s=d_s[iE], a0=d_aE[iE], t=dev(tE), noise=dev(nE), rate=1.0),
                 dict(net=m.net, sign=-1.0, s=d_s[iA], a0=d_aA[iA], t=dev(tA), noise=dev(nA), rate=1.0)]
        log.append(torch.stack(train_step(terms, opt, max_norm=1.0, precision=prec)))
    curve = torch.stack(log).cpu().numpy()
    ref = g["curve"]
    rel = np.abs(curve - ref) / np.maximum(1.0, np.abs(ref))
    k = 50
    ma = lambda x: np.stack([np.convolve(x[:, j], np.ones(k) / k, mode="valid") for j in range(2)], 1)
    mrel = np.abs(ma(curve) - ma(ref)) / np.abs(ma(ref))
    print(os.path.basename(path), prec, "first-50 max rel", rel[:50].max(), "| per-step max rel by 100s",
          [float(f"{rel[i:i + 100].max():.3g}") for i in range(0, steps, 100)], "| 50-step MA max rel", mrel.max(),
          "by 100s", [float(f"{mrel[i:i + 100].max():.3g}") for i in range(0, len(mrel), 100)],
          "| final losses", curve[-1], ref[-1], flush=True)
