"""CPU: the JSON contract of bench.py's reference arm (`--impl reference`) and the FLOP / byte bookkeeping the GPU arm's
roofline uses (SURVEY 8(d) figures)."""
import argparse
import json

import pytest

import bench


def test_reference_arm_line_has_the_contract_keys(monkeypatch, capsys):
    monkeypatch.setattr(bench, "cpu_port_rate", lambda task, B, n_steps, flat, threads, reps=1: (2000.0, B / 2000.0))
    monkeypatch.delenv("RANK", raising=False)
    args = argparse.Namespace(gpus=1, steps=2, warmup=1, impl="reference", workload="maze", precision="fp16", batch=0, no_cpu=False)
    bench.run_reference(args)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "scored_samples_per_sec" and line["unit"] == "samples/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1
    assert line["value"] == pytest.approx(2000.0) and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "sample" in line["cpu_baseline"]
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("maze: s 6 + a 2, 20-step ODE log-density")
    assert line["config"]["batch_per_gpu"] == 65536 and line["gpu_launches"] == 0


def test_reference_arm_other_ranks_print_nothing(monkeypatch, capsys):
    monkeypatch.setenv("RANK", "3")
    bench.run_reference(argparse.Namespace(gpus=8, steps=1, warmup=0, workload="maze"))
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
