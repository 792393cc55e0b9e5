"""CPU: the returns / GAE oracle (oracle/ppo_closed_form.py) against outputs of the reference's own
RolloutStorage.compute_returns / compute_advantages (tests/golden/gae_*.npz)."""
import glob
import os

import numpy as np
import pytest

from oracle import ppo_closed_form as ppo
from oracle.gen_golden import gae_inputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "gae_*.npz")))


def test_fixtures_present():
    assert len(CASES) >= 5


@pytest.mark.parametrize("path", CASES, ids=os.path.basename)
def test_oracle_is_bit_identical_to_reference_scan(path):
    g = np.load(path)
    T, N, D, use_gae, proper, seed = (int(x) for x in g["meta"])
    rewards, value_preds, masks, bad_masks, next_value = gae_inputs(seed, T, N, D)
    returns, vp = ppo.compute_returns(rewards, value_preds, masks, bad_masks, next_value, float(g["gamma"]),
                                      float(g["gae_lambda"]), bool(use_gae), bool(proper))
    assert np.array_equal(returns, g["returns"])
    assert np.array_equal(vp, g["value_preds_after"])
    adv = ppo.compute_advantages(returns, vp)
    assert np.abs(adv - g["advantages"]).max() <= 2e-6 * max(1.0, np.abs(g["advantages"]).max())
