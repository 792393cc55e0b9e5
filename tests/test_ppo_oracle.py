"""CPU: the actor-critic / clipped-PPO oracle (oracle/ppo_closed_form.py) against outputs of the reference's own MLPBasic
and DiagGaussian classes (tests/golden/ppo_*.npz, ppotrain_*.npz, written by oracle/gen_golden.py::gen_ppo)."""
import glob
import os

import numpy as np
import pytest

from oracle import ppo_closed_form as ppo
from oracle import synth
from oracle.gen_golden import ppo_inputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "ppo_*.npz")))


def load_case(path):
    g = np.load(path)
    s_dim, a_dim, hidden, layers, B, seed = (int(x) for x in g["meta"])
    spec = ppo.ACSpec(s_dim, a_dim, hidden, layers)
    flat = ppo.ac_make_params(spec, seed)
    return g, spec, flat, ppo_inputs(spec, flat, seed, B)


def test_fixtures_present():
    assert len(CASES) >= 5 and len(glob.glob(os.path.join(GOLDEN, "ppotrain_*.npz"))) >= 2


@pytest.mark.parametrize("path", CASES, ids=os.path.basename)
def test_oracle_matches_reference_policy_and_loss(path):
    g, spec, flat, (obs, action, old_logp, adv, ret, old_value, _) = load_case(path)
    clip, vcoef, ecoef = (float(x) for x in g["coefs"])
    value, logp, ent = ppo.ac_evaluate(spec, flat, obs, action)
    _, mean, _ = ppo.ac_forward(spec, flat, obs)
    for name, x in (("value", value), ("logp", logp), ("ent", ent), ("mean", mean)):
        assert np.abs(x - g[f"{name}_f64"]).max() <= 1e-11 * max(1.0, np.abs(x).max()), name
        assert np.abs(x - g[f"{name}_f32"]).max() <= 2e-5 * max(1.0, np.abs(x).max()), name
    losses, grad = ppo.ppo_loss_and_grad(spec, flat, obs, action, old_logp, adv, ret, old_value, clip, vcoef, ecoef)
    assert np.allclose(losses, g["loss_f64"], rtol=1e-11, atol=1e-13)
    assert np.allclose(losses, g["loss_f32"], rtol=1e-5, atol=2e-7)
    ref = g["grad_f64"]
    tol = 1e-12 if ref.dtype == np.float64 else 2e-7
    assert np.abs(grad - ref).max() <= tol * max(1.0, np.abs(ref).max())
    ref32 = g["grad_f32"].astype(np.float64)
    assert np.sqrt(((grad - ref32) ** 2).sum()) <= 1e-4 * np.sqrt((ref32 ** 2).sum())


def test_clip_branches_are_exercised():
    fr = np.array([np.load(p)["clip_frac"] for p in CASES])
    assert fr[:, 0].max() > 0.5 and fr[:, 1].max() > 0.2 and fr[:, 0].min() == 0.0


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ppotrain_*.npz"))), ids=os.path.basename)
def test_oracle_updates_follow_reference_curve(path):
    g = np.load(path)
    s_dim, a_dim, hidden, layers, rows, nmb, epochs, seed = (int(x) for x in g["meta"])
    spec = ppo.ACSpec(s_dim, a_dim, hidden, layers)
    flat0 = ppo.ac_make_params(spec, seed)
    data = ppo_inputs(spec, flat0, seed, rows)[:6]
    flat = flat0.astype(np.float64)
    m, v = np.zeros_like(flat), np.zeros_like(flat)
    mb, it = rows // nmb, 0
    for e in range(min(epochs, 5)):
        perm = synth.rng(seed + 100 + e).permutation(rows)
        for k in range(nmb):
            i = perm[k * mb:(k + 1) * mb]
            flat, m, v, losses, norm = ppo.ppo_step(spec, flat, m, v, it + 1, tuple(x[i] for x in data), float(g["lr"]))
            tol = 1e-4 if it < 8 else 3e-3
            assert np.allclose(losses, g["curve"][it], rtol=tol, atol=tol * 0.1), (it, losses, g["curve"][it])
            assert abs(norm - g["gradnorm"][it]) <= (2e-3 if it < 8 else 2e-2) * g["gradnorm"][it]
            it += 1


def test_oracle_limits():
    """Closed-form limits of the PPO oracle: on-policy data (ratio = 1) gives actor_loss = -mean(adv); a huge clip
    switches both clips off; GAE with lambda = 1 and no episode ends is the discounted return."""
    spec = ppo.ACSpec(7, 3, 32, 2)
    flat = ppo.ac_make_params(spec, 8).astype(np.float64)
    g = np.random.default_rng(1)
    B = 100
    obs, act = g.standard_normal((B, 7)), g.standard_normal((B, 3)) * 0.4
    value, logp, ent = ppo.ac_evaluate(spec, flat, obs, act)
    adv, ret = g.standard_normal((B, 1)), g.standard_normal((B, 1))
    losses, _ = ppo.ppo_loss_and_grad(spec, flat, obs, act, logp, adv, ret, value)
    assert abs(losses[2] + adv.mean()) < 1e-12 and abs(losses[1] - 0.5 * ((value - ret) ** 2).mean()) < 1e-12
    assert abs(losses[3] - ent.mean()) < 1e-12
    old_logp, old_value = logp + g.standard_normal((B, 1)), value + g.standard_normal((B, 1))
    losses, _ = ppo.ppo_loss_and_grad(spec, flat, obs, act, old_logp, adv, ret, old_value, clip=1e9)
    assert abs(losses[2] + (np.exp(logp - old_logp) * adv).mean()) < 1e-12
    assert abs(losses[1] - 0.5 * ((value - ret) ** 2).mean()) < 1e-12
    T, N, gamma = 12, 3, 0.9
    r = g.standard_normal((T, N, 1)).astype(np.float32)
    vp = g.standard_normal((T + 1, N, 1)).astype(np.float32)
    ones = np.ones((T + 1, N, 1), np.float32)
    nv = g.standard_normal((N, 1)).astype(np.float32)
    returns, _ = ppo.compute_returns(r, vp, ones, ones, nv, gamma, 1.0, True, False)
    want = np.zeros((T, N, 1))
    acc = nv.astype(np.float64)
    for t in reversed(range(T)):
        acc = r[t] + gamma * acc
        want[t] = acc
    assert np.abs(returns[:-1] - want).max() < 1e-4
