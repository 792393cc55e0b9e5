"""CPU: the JSON contract of bench.py's reference arm (`--impl reference`) and the FLOP / byte bookkeeping the GPU arm's
roofline uses (SURVEY 8(d) figures)."""
import argparse
import json

import pytest

import bench


def _fake_arm(monkeypatch, kind="port"):
    class Arm:
        s_dim, a_dim = 6, 2
        desc = "stand-in"

        def __init__(self, task, flat=None):
            self.kind = kind

        def score_rate(self, B, n_steps, threads, reps=1):
            return 2000.0, B / 2000.0

        def train_rate(self, B, threads, steps=3):
            return 50000.0, 2 * B / 50000.0

    monkeypatch.setattr(bench, "CpuArm", Arm)


def test_reference_arm_line_has_the_contract_keys(monkeypatch, capsys):
    _fake_arm(monkeypatch)
    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    args = argparse.Namespace(gpus=1, steps=2, warmup=1, impl="reference", workload="maze", precision="fp16x3", batch=0,
                              no_cpu=False, metric="score")
    bench.run_reference(args)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "scored_samples_per_sec" and line["unit"] == "samples/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1
    assert line["value"] == pytest.approx(2000.0) and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "sample" in line["cpu_baseline"]
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("maze: s 6 + a 2, 20-step ODE log-density")
    assert line["config"]["batch_per_gpu"] == 65536 and line["gpu_launches"] == 0
    # both arms build `config` with the same function: identical dictionaries for the same workload and GPU count
    assert line["config"] == bench.config_dict("maze", 6, 2, 20, 65536, 1)


def test_reference_arm_train_metric(monkeypatch, capsys):
    _fake_arm(monkeypatch, "reference")
    monkeypatch.delenv("RANK", raising=False)
    args = argparse.Namespace(gpus=1, steps=1, warmup=0, impl="reference", workload="maze", precision="fp16x3", batch=0,
                              no_cpu=False, metric="train")
    bench.run_reference(args)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["metric"] == "cfm_train_samples_per_sec" and line["cpu_baseline"]["kind"] == "reference"
    assert line["value"] == pytest.approx(50000.0, rel=1e-6)


def test_cpu_arm_times_the_staged_reference_when_present():
    """oracle/_ref (built by oracle/build_ref.py from the reference checkout) makes the CPU arm the reference's own classes;
    its log-density agrees with the oracle on the arm's own weights."""
    import numpy as np
    import torch

    from oracle import build_ref
    from oracle import closed_form as cf

    if build_ref.build() is None:
        pytest.skip("no reference checkout and no staged oracle/_ref")
    arm = bench.CpuArm("maze", cf.make_params(cf.NetSpec(6, 2, 256, 4, 2), 3, 1.0))
    assert arm.kind == "reference"
    rate, dt = arm.score_rate(64, 4, 1)
    assert rate > 0 and dt > 0
    _, _, s, a, probe = bench.make_inputs("maze", 64, 4)
    ref = cf.log_prob(arm.spec, arm.flat, s, a, "expert", np.ones_like(probe), 4, exact=True)   # a_dim = 2: probe-free check
    with torch.enable_grad():
        runs = []
        for i in range(2):   # exact trace from two basis-probe runs of the reference (as oracle/gen_golden.py does)
            e = torch.zeros(64, 2)
            e[:, i] = 1.0
            arm.P._rademacher, saved = (lambda *a_, **k: e.clone()), arm.P._rademacher
            try:
                runs.append(arm.model.log_prob(torch.from_numpy(s), torch.from_numpy(a), "expert", n_steps=4).detach().numpy())
            finally:
                arm.P._rademacher = saved
        z = torch.zeros(64, 2)
        arm.P._rademacher, saved = (lambda *a_, **k: z.clone()), arm.P._rademacher
        try:
            zero = arm.model.log_prob(torch.from_numpy(s), torch.from_numpy(a), "expert", n_steps=4).detach().numpy()
        finally:
            arm.P._rademacher = saved
    got = runs[0].astype(np.float64) + runs[1] - zero
    assert np.abs(got - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())


def test_reference_arm_other_ranks_print_nothing(monkeypatch, capsys):
    monkeypatch.setenv("RANK", "3")
    bench.run_reference(argparse.Namespace(gpus=8, steps=1, warmup=0, workload="maze", metric="score", batch=0))
    assert capsys.readouterr().out == ""


def test_algorithmic_flops_match_the_survey_table():
    # scored sample @32 steps = 51.0 (maze 6+2), 51.3 (hopper), 52.1 (cheetah), 56.3 (ant 132+8), 56.5 (hand) MFLOP;
    # maze @20 steps = 31.8 MFLOP
    f = bench.algorithmic_flops_per_sample
    assert f(6, 2, 20) == 31846400
    for (s, a), mflop in (((6, 2), 51.0), ((11, 3), 51.3), ((17, 6), 52.1), ((132, 8), 56.3), ((68, 20), 56.5)):
        assert abs(f(s, a, 32) / 1e6 - mflop) < 0.06, (s, a, f(s, a, 32))


def test_every_baseline_config_has_a_workload():
    for name in ("sine", "maze", "hopper", "halfcheetah", "walker", "ant", "hand", "pick", "push"):
        task, B, n_steps = bench.WORKLOADS[name]
        assert B >= 65536 and n_steps in (20, 32)
