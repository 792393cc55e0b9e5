"""CPU oracle for the returns / GAE recursion and the advantage normalisation of the PPO consumer (scope row f4, first
part): numpy float32 restatement of ``RolloutStorage.compute_returns`` / ``compute_advantages``
(deps/rl-toolkit/rlf/storage/rollout_storage.py:178-239), one numpy op per torch op so the scan is bit-identical.

TEST INFRASTRUCTURE ONLY: nothing under ``modril_b200/`` imports this module.  Pinned to outputs of the reference's own
two methods by ``oracle/gen_golden.py::gen_gae`` (``tests/golden/gae_*.npz``).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def compute_returns(rewards, value_preds, masks, bad_masks, next_value, gamma, gae_lambda, use_gae,
                    use_proper_time_limits):
    """rewards [T,N,1], value_preds [T+1,N,D], masks / bad_masks [T+1,N,1], next_value [N,D] -> (returns [T+1,N,D],
    value_preds after the call).  Entries of ``returns`` the reference leaves untouched (returns[T] under GAE) are 0."""
    rewards, masks, bad_masks = (np.asarray(x, f32) for x in (rewards, masks, bad_masks))
    value_preds = np.array(value_preds, f32)
    T, D = rewards.shape[0], value_preds.shape[2]
    exp_rewards = np.repeat(rewards, D, axis=2)                       # :179
    returns = np.zeros_like(value_preds)
    g, gl = f32(gamma), f32(gamma * gae_lambda)
    if use_gae:
        value_preds[-1] = next_value                                   # :185 / :210
        gae = np.zeros_like(value_preds[0])
        for step in reversed(range(T)):
            delta = exp_rewards[step] + g * value_preds[step + 1] * masks[step + 1] - value_preds[step]
            gae = delta + gl * masks[step + 1] * gae
            if use_proper_time_limits:
                gae = gae * bad_masks[step + 1]                        # :198
            returns[step] = gae + value_preds[step]
    else:
        returns[-1] = next_value                                       # :201 / :226
        for step in reversed(range(T)):
            r = returns[step + 1] * g * masks[step + 1] + exp_rewards[step]
            if use_proper_time_limits:                                 # :205-211
                r = r * bad_masks[step + 1] + (f32(1) - bad_masks[step + 1]) * value_preds[step]
            returns[step] = r
    return returns, value_preds


def compute_advantages(returns, value_preds, eps=1e-5):
    """:235-239 -- (adv - mean) / (std + 1e-5) with the unbiased std; moments in float64 (torch's float32 reductions
    agree to an ulp or two, the test tolerance says so)."""
    adv = np.asarray(returns, f32)[:-1] - np.asarray(value_preds, f32)[:-1]
    mean = f32(adv.astype(np.float64).mean())
    std = f32(adv.astype(np.float64).std(ddof=1))
    return (adv - mean) / (std + f32(eps))
