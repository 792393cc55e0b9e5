"""CPU: the oracle's update (oracle/closed_form.py::train_step, float64, hand-written backward) replays the reference's
offline pre-training loop (tests/golden/pretrain_*.npz: pretrain.py:269-440 run on synthetic demonstrations)."""
import glob
import os

import numpy as np
import pytest

from oracle import closed_form as cf
from oracle import synth
from oracle.gen_golden import pretrain_dataset, pretrain_draws

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "pretrain_*.npz")))


def test_fixtures_present():
    assert len(CASES) >= 3


@pytest.mark.parametrize("path", CASES, ids=os.path.basename)
def test_oracle_replays_reference_pretraining(path):
    g = np.load(path)
    s_dim, a_dim, hidden, rows, epochs, seed, n_heads, do_norm = (int(x) for x in g["meta"])
    mode = ["autotune", "fixed", "2fs"][int(g["mode"])]
    # data preparation (pretrain.py:33-39, :273-290) restated in numpy
    task = os.path.basename(path).split("_")[2]
    obs, act = pretrain_dataset(seed, task, rows)
    if do_norm:
        nv = lambda x: np.clip((x - x.mean(0)) / (x.std(0, ddof=1) + 1e-8), -10.0, 10.0)
        obs, act = nv(obs.astype(np.float64)), nv(act.astype(np.float64))
    data = np.concatenate([obs, act], 1)
    if data.shape[0] % 2 == 1:
        data = data[1:]
    assert np.abs(data - g["dataset"]).max() <= 2e-5
    data = g["dataset"]
    spec = cf.NetSpec(s_dim, a_dim, hidden, 4, n_heads)
    flat = cf.make_params(spec, seed + 5, 1.0).astype(np.float64)
    m, v = np.zeros_like(flat), np.zeros_like(flat)
    n, B, it = data.shape[0], 128, 0
    for e in range(1):                                     # the first epoch is enough to pin the composition
        perm = synth.rng(seed + 50 + e).permutation(n)
        for k, lo in enumerate(range(0, n, B)):
            i = perm[lo:lo + B]
            s, a = data[i, :s_dim], data[i, s_dim:]
            d = pretrain_draws(seed, e, k, len(i), a_dim, 1e-2)
            if mode == "2fs":
                a0 = (a + d["z"] * np.float32(0.02)).astype(np.float32)
                flat, m, v, losses, norm = cf.train_step(spec, flat, m, v, it + 1, [("expert", s, a0, d["tE"], d["nE"])],
                                                         float(g["lr"]), 1.0, rates=(1.0,))
                total = losses[0]
            else:
                a_noise = (a + np.float32(0.1) * d["z"]).astype(np.float32)
                sp = ("warmup", 1e-3, 1e-8) if mode == "autotune" else None
                reg = dict(s=s, a=a_noise, t=d["t_mix"], probe=d["probe"], anti=sp or 1e-4, stable=sp or 1e-6)
                flat, m, v, losses, norm, (la, ls, wa, ws) = cf.train_step(
                    spec, flat, m, v, it + 1, [("expert", s, a, d["tE"], d["nE"]), ("agent", s, a_noise, d["tA"], d["nA"])],
                    float(g["lr"]), 1.0, reg=reg)
                total = losses[0] + losses[1] + wa * la + ws * ls
                assert np.allclose([wa, ws], g["weights"][it], rtol=2e-2), (it, wa, ws, g["weights"][it])
            ref = g["curve"][it]
            assert abs(total - ref[0]) <= 3e-3 * abs(ref[0]) and abs(losses[0] - ref[1]) <= 3e-3 * abs(ref[1]), (it, total, ref)
            assert abs(norm - g["gradnorm"][it]) <= 1e-2 * g["gradnorm"][it]
            it += 1
