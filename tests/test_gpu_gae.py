"""GPU parity of frl_compute_returns / frl_normalise_advantages (through the C ABI) against the reference goldens."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import ppo_closed_form as ppo
from oracle.gen_golden import gae_inputs
from tests.gpu_common import dev

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run(T, N, D, gamma, lam, use_gae, proper, inputs):
    from modril_b200.ppo import compute_advantages, compute_returns

    rewards, value_preds, masks, bad_masks, next_value = (dev(x) for x in inputs)
    returns = torch.zeros(T + 1, N, D, device="cuda:0")
    compute_returns(rewards, value_preds, masks, bad_masks if proper else None, next_value, returns, gamma=gamma,
                    gae_lambda=lam, use_gae=use_gae, use_proper_time_limits=proper)
    adv, stats = compute_advantages(returns, value_preds, return_stats=True)
    return returns.cpu().numpy(), value_preds.cpu().numpy(), adv.cpu().numpy(), stats.cpu().numpy()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "gae_*.npz"))), ids=os.path.basename)
def test_returns_bit_identical_and_advantages_close(path):
    g = np.load(path)
    T, N, D, use_gae, proper, seed = (int(x) for x in g["meta"])
    returns, vp, adv, stats = _run(T, N, D, float(g["gamma"]), float(g["gae_lambda"]), bool(use_gae), bool(proper),
                                   gae_inputs(seed, T, N, D))
    assert np.array_equal(returns, g["returns"])
    assert np.array_equal(vp, g["value_preds_after"])
    assert np.abs(adv - g["advantages"]).max() <= 2e-6 * max(1.0, np.abs(g["advantages"]).max())
    raw = g["returns"][:-1] - g["value_preds_after"][:-1]
    assert abs(stats[0] - raw.mean(dtype=np.float64)) <= 1e-6 and abs(stats[1] - raw.std(ddof=1, dtype=np.float64)) <= 1e-6


@pytest.mark.parametrize("use_gae,proper", [(True, True), (False, False)])
def test_large_rollout_matches_oracle(use_gae, proper):
    T, N, D = 256, 4099, 1
    inputs = gae_inputs(77, T, N, D)
    returns, vp, adv, _ = _run(T, N, D, 0.995, 0.97, use_gae, proper, inputs)
    r_ref, vp_ref = ppo.compute_returns(*inputs, 0.995, 0.97, use_gae, proper)
    assert np.array_equal(returns, r_ref) and np.array_equal(vp, vp_ref)
    a_ref = ppo.compute_advantages(r_ref, vp_ref)
    assert np.abs(adv - a_ref).max() <= 2e-6 * max(1.0, np.abs(a_ref).max())
    assert abs(float(adv.mean(dtype=np.float64))) < 1e-6 and abs(float(adv.std(ddof=1, dtype=np.float64)) - 1) < 1e-4


def test_bad_arguments_raise():
    from modril_b200.ppo import compute_advantages, compute_returns

    z = lambda *s: torch.zeros(*s, device="cuda:0")
    with pytest.raises(ValueError):
        compute_returns(z(4, 2, 1), z(5, 2, 1), z(5, 2, 1), None, z(2, 1), z(5, 2, 1), gamma=0.99, use_proper_time_limits=True)
    with pytest.raises(ValueError):
        compute_returns(z(4, 2, 1), z(4, 2, 1), z(5, 2, 1), None, z(2, 1), z(5, 2, 1), gamma=0.99)
    with pytest.raises(RuntimeError):
        compute_advantages(torch.zeros(5, 2, 1), torch.zeros(5, 2, 1))
