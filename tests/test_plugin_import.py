"""CPU: with rl-toolkit importable (the staged, unmodified slice under oracle/_ref/rl-toolkit) the reward module derives from
the REAL plugin classes -- ``FlowMatchingEstimation(rlf.algos.il.base_irl.BaseIRLAlgo)``, ``FLOWRIL(rlf.algos.nested_algo
.NestedAlgo)`` -- inherits rl-toolkit's own ``update`` loop instead of the stand-in, and its flags parse through the real
``get_add_args`` chain (BaseAlgo -> BaseNetAlgo -> BaseILAlgo -> module).  Runs in a fresh interpreter: the base class is
chosen when ``modril_b200.flowril.flowril`` is imported."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_runner(*argv, timeout=300):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "plugin_loop_runner.py"), *argv], capture_output=True,
                         text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-4000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    if "skipped" in res:
        pytest.skip(res["skipped"])
    return res


def test_module_plugs_into_the_real_rl_toolkit_classes():
    res = run_runner("import")
    assert res["have_rlf"] and res["module_is_rlf_subclass"] and res["flowril_is_nested_algo"]
    assert res["update_is_the_reference_loop"]      # base_irl.py:37-72 itself, not the stand-in
    for flag in ("flowril_state_norm", "flowril_reward_norm", "flowril_precision", "flowril_host_rewards", "option",
                 "hidden_dim", "disc_lr", "reward_type"):
        assert flag in res["flags"]
