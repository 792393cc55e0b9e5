"""Drives the reward module through rl-toolkit's OWN update loop (run as a subprocess by tests/test_gpu_plugin_loop.py and
tests/test_plugin_import.py; a fresh interpreter because ``modril_b200.flowril.flowril`` decides at import time whether it
derives from the real ``rlf`` classes).

    python tests/plugin_loop_runner.py import        # CPU: import chain + flag parsing only
    python tests/plugin_loop_runner.py loop [2fs]    # cuda:0: NestedAlgo.update x 3 over a real RolloutStorage

The caller side is the reference's code, unmodified, staged by oracle/build_ref.py (oracle/_ref/rl-toolkit):
``NestedAlgo.update`` (rlf/algos/nested_algo.py:91-106) -> ``BaseIRLAlgo.update`` (rlf/algos/il/base_irl.py:37-72) ->
our ``_update_reward_func`` / ``_get_reward`` hooks, with rlf's ``RolloutStorage`` and ``TransitionDataset``.  Prints one
JSON line.
"""
import argparse
import json
import os
import sys
import tempfile
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import build_ref  # noqa: E402

RLF = build_ref.rlf_path()
if RLF is None:
    print(json.dumps({"skipped": "oracle/_ref/rl-toolkit is not staged (python -m oracle.build_ref in the build container)"}))
    sys.exit(0)
build_ref.install_sim_stubs()
sys.path.insert(0, RLF)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def make_args(algo, extra):
    parser = argparse.ArgumentParser(conflict_handler="resolve")
    algo.get_add_args(parser)
    args = parser.parse_args(extra)
    # what rl-toolkit's runner adds (rlf/args.py) and the loop reads
    for k, v in dict(num_steps=16, num_processes=8, log_smooth_len=10, recurrent_policy=False, policy_ob_key="observation",
                     gamma=0.99, gae_lambda=0.95, use_gae=True, use_proper_time_limits=False, num_env_steps=16 * 8 * 100,
                     seed=31, freeze_policy=False).items():
        setattr(args, k, v)
    return args


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "import"
    option = sys.argv[2] if len(sys.argv) > 2 else "scrf"
    import gym
    import rlf.algos.il.base_irl as base_irl
    import rlf.algos.nested_algo as nested
    from rlf.storage.rollout_storage import RolloutStorage

    from modril_b200.flowril import flowril as fl
    from modril_b200.ppo import PPO as DevicePPO

    out = {"have_rlf": bool(fl.HAVE_RLF),
           "module_is_rlf_subclass": issubclass(fl.FlowMatchingEstimation, base_irl.BaseIRLAlgo),
           "flowril_is_nested_algo": issubclass(fl.FLOWRIL, nested.NestedAlgo),
           "update_is_the_reference_loop": fl.FlowMatchingEstimation.update is base_irl.BaseIRLAlgo.update}
    algo = fl.FLOWRIL(agent_updater=DevicePPO())
    S, A = 11, 3
    tmp = tempfile.mkdtemp()
    args = make_args(algo, ["--traj-load-path", "demos.pt", "--traj-batch-size", "32", "--option", option, "--hidden-dim", "128",
                            "--linear-lr-decay", "False"])
    args.cwd = tmp
    out["flags"] = sorted(k for k in vars(args) if k.startswith("flowril") or k in ("option", "hidden_dim", "disc_lr", "reward_type"))
    if mode == "import":
        print(json.dumps(out))
        return

    from modril_b200.ppo import DistActorCritic

    dev = torch.device("cuda", 0)
    args.device = dev
    g = torch.Generator().manual_seed(0)
    n_demo = 256
    torch.save({"obs": torch.randn(n_demo, S, generator=g), "next_obs": torch.randn(n_demo, S, generator=g),
                "actions": torch.tanh(torch.randn(n_demo, A, generator=g)),
                "done": (torch.arange(n_demo) % 64 == 63).float()}, os.path.join(tmp, "demos.pt"))
    obs_space = gym.spaces.Box(low=-np.inf, high=np.inf, shape=(S,), dtype=np.float32)
    ac_space = gym.spaces.Box(low=-1.0, high=1.0, shape=(A,), dtype=np.float32)
    policy = DistActorCritic(S, A, 64, 2).to(dev)
    policy.obs_space, policy.action_space = obs_space, ac_space
    algo.init(policy, args)
    algo.set_env_ref(None)          # no VecNormalize: get_env_ob_filt() -> None (rlf/algos/base_algo.py:66-83)
    T, N = args.num_steps, args.num_processes
    storage = RolloutStorage(T, N, obs_space, ac_space, args)
    storage.to(dev)
    module = algo.modules[0]

    # count the device->host synchronisations made inside OUR hooks (torch's sync debug mode warns on each one)
    # ... attributed to the Python line that triggered them: `ours` = lines of modril_b200/, `caller` = rl-toolkit's own
    # code running inside the hook (the rollout generator indexes seven device tensors with a Python list per minibatch)
    syncs = {"update_reward_func": 0, "get_reward": 0, "caller": 0}
    where = {}

    def counted(name, fn):
        def wrapper(*a, **k):
            with warnings.catch_warnings(record=True) as w:
                warnings.simplefilter("always")
                torch.cuda.set_sync_debug_mode("warn")
                try:
                    return fn(*a, **k)
                finally:
                    torch.cuda.set_sync_debug_mode("default")
                    for x in w:
                        if "synchroniz" not in str(x.message).lower():
                            continue
                        ours = os.sep + "modril_b200" + os.sep in x.filename
                        syncs[name if ours else "caller"] += 1
                        key = f"{os.path.relpath(x.filename, ROOT)}:{x.lineno}"
                        where[key] = where.get(key, 0) + 1
        return wrapper

    module._update_reward_func = counted("update_reward_func", module._update_reward_func)
    module._get_reward = counted("get_reward", module._get_reward)

    logs = []
    for it in range(3):
        storage.obs.copy_(torch.randn(T + 1, N, S, generator=g))
        storage.actions.copy_(torch.tanh(torch.randn(T, N, A, generator=g)))
        masks = torch.ones(T + 1, N, 1)
        masks[5 + it, 2] = 0.0          # an episode of process 2 ends inside the rollout
        masks[T - 1, 5] = 0.0
        storage.masks.copy_(masks)
        with torch.no_grad():
            pol = policy.get_action(storage.obs[:-1].reshape(T * N, S))
        storage.value_preds[:-1].copy_(pol["value"].reshape(T, N, 1))
        storage.action_log_probs.copy_(pol["action_log_probs"].reshape(T, N, 1))
        storage.rewards.fill_(123.0)     # "environment rewards": must be wiped and replaced by the IRL reward
        for k in syncs:
            syncs[k] = 0
        where.clear()
        algo.pre_update(it)
        log = algo.update(storage, args, False, 1)
        rewards = storage.rewards.detach().cpu()
        # the same rollout scored directly (no loop): what the reward block must equal before normalisation is not
        # observable from outside, so check the invariants instead: finite, not the wiped value, one row per step
        logs.append({"keys": sorted(log.keys()), "culm_irl_reward": float(log.get("culm_irl_reward", float("nan"))),
                     "flow_loss": float(log.get("flow_loss", float("nan"))), "agent_loss": float(log.get("agent_loss", float("nan"))),
                     "rewards_finite": bool(torch.isfinite(rewards).all()), "rewards_replaced": bool((rewards != 123.0).all()),
                     "rewards_std": float(rewards.std()), "syncs": dict(syncs),
                     "episodes_logged": len(module.ep_log_vals["reward"])})
    out["updates"] = logs
    out["sync_sites_last_update"] = where
    # the per-step rows the real loop wrote == the module's own batched block (host copy round trip is exact)
    block = module._norm_cache if module._norm_cache is not None else module._rollout_cache
    out["storage_equals_block"] = bool(torch.equal(storage.rewards.detach().cpu(), block.detach().cpu()))
    out["reward_rows_on_host"] = bool(module._serve is not None and module._serve.device.type == "cpu")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
