"""LIVE check of the oracle and of the committed fixtures against the reference itself (build container only; skipped where
``/root/reference`` is absent, e.g. on the GPU box).  A subset of the goldens is regenerated on the spot by importing the
UNMODIFIED reference (oracle/ref_shim.py) and running the same helper functions oracle/gen_golden.py used:
  (i)  the reference's output today == the committed fixture, bit for bit (the fixtures are what the reference computes);
  (ii) the numpy oracle agrees with the live reference at the bar tests/test_oracle_vs_golden.py uses on the fixtures.
Covers log-density (scrf shared probe / 2fs per-step probe), CFM loss + gradient, the return normalisation (row a10: the
reference's own RunningMeanStd) and the op-for-op torch port used as the fallback CPU baseline."""
import glob
import os

import numpy as np
import pytest

from oracle import closed_form as cf
from oracle import ref_shim, synth

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref():
    import torch

    torch.set_num_threads(4)
    from oracle import gen_golden as gg

    return ref_shim.load(), gg


def _rel(got, want):
    return np.abs(np.asarray(got, np.float64) - want) / np.maximum(1.0, np.abs(want))


@pytest.mark.parametrize("pattern", ["logprob_00_*", "logprob_0[3-4]_*", "logprob_s00_*"])
def test_logprob_fixture_is_the_reference_output_and_the_oracle_follows_it(ref, pattern):
    P, gg = ref
    paths = sorted(glob.glob(os.path.join(GOLDEN, pattern + ".npz")))
    assert paths
    for path in paths[:2]:
        g = np.load(path)
        s_dim, a_dim, hidden, depth, n_heads, n_steps, B, seed = (int(x) for x in g["meta"])
        spec = cf.NetSpec(s_dim, a_dim, hidden, depth, n_heads)
        gain = float(g["gain"])
        s, a = synth.states_actions(seed, B, s_dim, a_dim)
        probe = synth.rademacher(seed + 1, (B, a_dim))
        probe_steps = synth.rademacher(seed + 2, (n_steps, B, a_dim))
        for ri, role in enumerate(("expert", "agent")):
            flat = cf.make_params(spec, seed + (10 * ri if n_heads == 1 else 0), gain)
            model = gg.build_ref(P, spec, flat)
            for key, pr in ((f"logp_{role}", probe), (f"logp_{role}_stepprobe", probe_steps)):
                live = gg.ref_log_prob(P, model, spec, s, a, role, pr, n_steps)
                assert np.array_equal(live, g[key]), f"{os.path.basename(path)} {key}: the reference no longer reproduces the fixture"
                err = _rel(cf.log_prob(spec, flat, s, a, role, pr, n_steps), live.astype(np.float64))
                # fp32 reference vs float64 oracle: 2e-5, except rows on a ReLU kink (<= 1 of 96 here, <= 1e-3)
                assert np.sort(err)[-2] <= 2e-5 and err.max() <= 1e-3, (key, err.max())


@pytest.mark.parametrize("pattern", ["fmloss_00_*", "fmloss_01_*"])
def test_fmloss_fixture_is_the_reference_output_and_the_oracle_follows_it(ref, pattern):
    P, gg = ref
    path = sorted(glob.glob(os.path.join(GOLDEN, pattern + ".npz")))[0]
    g = np.load(path)
    s_dim, a_dim, hidden, depth, n_heads, _, B, seed = (int(x) for x in g["meta"])
    spec = cf.NetSpec(s_dim, a_dim, hidden, depth, n_heads)
    flat = cf.make_params(spec, seed, float(g["gain"]))
    s, a = synth.states_actions(seed, B, s_dim, a_dim)
    for ri, role in enumerate(("expert", "agent")):
        model = gg.build_ref(P, spec, flat)
        t, noise, u = synth.cfm_draws(seed + 1 + ri, B, a_dim, return_u=True)
        dq = synth.rng(seed + 5 + ri).standard_normal((B, a_dim)).astype(np.float32) if n_heads == 1 else None
        loss = gg.ref_fm_loss(P, model, spec, s, a, role, u, noise, dq)
        loss.backward()
        grad = gg.flat_grads(model)
        assert np.float32(loss.item()) == g[f"loss_{role}"]
        assert np.array_equal(grad, g[f"grad_{role}"])
        ref_loss, ref_grad = cf.fm_loss_and_grad(spec, flat, s, a, role, t, noise,
                                                 dequant=None if dq is None else dq * np.float32(0.02))
        assert abs(ref_loss - loss.item()) <= 2e-5 * max(1.0, abs(ref_loss))
        assert np.sqrt(((grad - ref_grad) ** 2).sum()) <= 1e-4 * np.sqrt((ref_grad ** 2).sum())


def test_return_normalisation_oracle_follows_the_reference_running_mean_std(ref):
    """Row a10 live: rlf/baselines/common/running_mean_std.py (imported by path) driven by the flowril.py:221-228 recursion
    == the oracle's restatement, rollout after rollout (carried returns and running moments included)."""
    import importlib.util

    rms_path = os.path.join(ref_shim.REFERENCE_ROOT, "deps", "rl-toolkit", "rlf", "baselines", "common", "running_mean_std.py")
    sp = importlib.util.spec_from_file_location("ref_running_mean_std_live", rms_path)
    mod = importlib.util.module_from_spec(sp)
    sp.loader.exec_module(mod)
    T, N, gamma = 16, 5, 0.99
    g = synth.rng(77)
    ref_rms, returns = mod.RunningMeanStd(shape=()), None
    ora_rms, ora_returns = None, None
    for rollout in range(3):
        rewards = (g.standard_normal((T, N, 1)) * 3.0 - 1.0).astype(np.float32)
        masks = (g.random((T, N, 1)) > 0.1).astype(np.float32)
        want = np.empty_like(rewards)
        for t in range(T):   # flowril.py:221-228 on torch float32 tensors, as the module runs it
            import torch

            reward, m = torch.from_numpy(rewards[t]), torch.from_numpy(masks[t])
            if returns is None:
                returns = reward.clone()
            returns = returns * m * gamma + reward
            ref_rms.update(returns.cpu().numpy())
            want[t] = (reward / np.sqrt(ref_rms.var[0] + 1e-8)).numpy()
        got, ora_returns, ora_rms = cf.normalise_rewards(rewards, masks, gamma, ora_rms, ora_returns)
        assert np.allclose(got, want, rtol=2e-6, atol=1e-7)
        assert np.allclose(ora_returns, returns.numpy(), rtol=2e-6, atol=1e-7)
        assert np.allclose(ora_rms.var, ref_rms.var, rtol=1e-6) and np.allclose(ora_rms.mean, ref_rms.mean, rtol=1e-6, atol=1e-7)


def test_torch_port_follows_the_live_reference(ref):
    """oracle/torch_port.py (the fallback CPU baseline of bench.py) == the reference's own log_prob and fm_loss."""
    import torch

    from oracle import torch_port as tp

    P, gg = ref
    spec = cf.NetSpec(11, 3, 128, 4, 2)
    flat = cf.make_params(spec, 5, 2.0)
    model = gg.build_ref(P, spec, flat)
    s, a = synth.states_actions(6, 64, 11, 3)
    probe = synth.rademacher(7, (64, 3))
    for role in ("expert", "agent"):
        live = gg.ref_log_prob(P, model, spec, s, a, role, probe, 16)
        got = tp.log_prob(spec, torch.from_numpy(flat), torch.from_numpy(s), torch.from_numpy(a), role, torch.from_numpy(probe), 16)
        assert np.allclose(got.detach().numpy(), live, rtol=1e-5, atol=1e-5)
    t, noise, u = synth.cfm_draws(8, 64, 3, return_u=True)
    live = gg.ref_fm_loss(P, model, spec, s, a, "expert", u, noise).item()
    got = tp.fm_loss(spec, torch.from_numpy(flat), torch.from_numpy(s), torch.from_numpy(a), "expert", torch.from_numpy(t),
                     torch.from_numpy(noise)).item()
    assert abs(got - live) <= 1e-5 * max(1.0, abs(live))
