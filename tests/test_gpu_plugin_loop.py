"""GPU: the reward module driven by rl-toolkit's OWN loop -- ``NestedAlgo.update`` (rlf/algos/nested_algo.py:91-106) ->
``BaseIRLAlgo.update`` (rlf/algos/il/base_irl.py:37-72) -> our hooks -> the device PPO updater -- over rlf's
``RolloutStorage`` and ``TransitionDataset`` (all unmodified reference code staged under oracle/_ref/rl-toolkit, see
oracle/build_ref.py).  Checks what the stand-in loop cannot: ``culm_irl_reward`` comes out of the reference's own logging
code, the environment rewards are wiped and replaced, and our hooks synchronise the host at most twice per update (the
logged losses; ONE copy of the [T, N] reward block) although the loop reads every reward with ``.item()``."""
import math

import pytest

from tests.test_plugin_import import run_runner

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(600)
@pytest.mark.parametrize("option", ["scrf", "2fs"])
def test_real_rl_toolkit_update_loop(option):
    res = run_runner("loop", option, timeout=540)
    assert res["have_rlf"] and res["update_is_the_reference_loop"]
    assert len(res["updates"]) == 3
    for i, u in enumerate(res["updates"]):
        assert "culm_irl_reward" in u["keys"] and "flow_loss" in u["keys"] and "agent_loss" in u["keys"], u["keys"]
        assert {"value_loss", "actor_loss", "dist_entropy"} <= set(u["keys"])   # NestedAlgo ran the agent updater after it
        assert math.isfinite(u["culm_irl_reward"]) and math.isfinite(u["flow_loss"]) and math.isfinite(u["agent_loss"])
        assert u["rewards_finite"] and u["rewards_replaced"] and u["rewards_std"] > 0
        assert u["episodes_logged"] == 2 * (i + 1)               # two episode ends per rollout reach ep_log_vals
        # host synchronisations triggered by lines of modril_b200/ inside the two hooks (torch's sync debug mode; the
        # rollout generator of rl-toolkit itself, which runs inside _update_reward_func, adds 7 per minibatch and is
        # counted separately as "caller"): ONE for the logged losses, ONE for the [T, N] reward block
        assert u["syncs"]["update_reward_func"] <= 1, u["syncs"]
        assert u["syncs"]["get_reward"] <= (1 if i > 0 else 4), u["syncs"]   # (first update: one-time probe-replay self-check)
    assert res["storage_equals_block"]
    assert res["reward_rows_on_host"]
