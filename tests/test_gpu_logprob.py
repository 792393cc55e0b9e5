"""GPU parity of the fused log-density integrator / velocity field / reward (through the C ABI).

Bars: the parity modes (fp32 FFMA kernels; fp16x3 = 3-term fp16 split on the tensor cores, the default) 1e-4 *
max(1,|ref|) per sample against (i) the committed outputs of the reference code and (ii) the float64 oracle; the
single-pass 16-bit tensor-core modes 2e-3 per sample on default-init-scale and trained nets (see TC_MODES below for what
they do NOT meet).  A ReLU-kink row is excused only where stated (gpu_common.assert_close)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import synth
from tests.gpu_common import assert_close, build_model, dev, spec_of

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Parity modes (bar 1e-4 per sample): "fp32" = FFMA kernels, "fp16x3" = the 3-term fp16 split on the tensor cores (the
# module default).  A golden has 96 rows; ONE row per output may sit on a ReLU kink (a pre-activation within rounding
# error of 0 flips a mask between two correct implementations: the reference's own fp32 result differs from the float64
# oracle on 3 of ~8000 golden rows by 3e-4 .. 6e-4).  0.011 = 1 / 96 rounded up; the rate over ALL goldens is pinned
# separately in test_parity_modes_kink_rate_over_all_goldens.
MODES = [("fp32", 1e-4, 0.011), ("fp16x3", 1e-4, 0.011)]
# Single-pass 16-bit tensor-core modes: north_star's bar is 2e-3 per sample.  It is asserted per sample -- no outlier
# allowance except rows on which the FP32 kernel itself misses 1e-4 (a kink in fp32 too), cap 5e-3 -- on nets at the scale
# of the reference's default init (the logprob_s* goldens, gain 1) and on CFM-trained nets (bench.py `parity`).  On the
# gain-2 stress goldens operand rounding (2^-11 fp16, 2^-8 bf16) flips ReLU masks on a large share of the rows and the
# single-pass modes DO NOT meet 2e-3 (fp16: 4-13 % of rows over, bf16: about half; README 'Precision modes'): there the
# tests only check that the kernels run deterministically and stay in the distribution measured in DESIGN.md section 5 --
# a regression guard, not a parity claim.  Parity on those nets is what the fp16x3 mode is for.
TC_MODES = {"fp16": dict(stress_median=1.5e-3, stress_cap=0.1), "bf16": dict(stress_median=6e-3, stress_cap=0.25)}


def _modes():
    return MODES


def _check_tc(got, ref, fp32_got, mode, gain, what):
    from tests.gpu_common import rel_err
    err = rel_err(got, ref)
    if gain <= 1.0:
        excused = rel_err(fp32_got, ref) > 1e-4          # the row is a kink for the FP32 kernel as well
        bad = (err > 2e-3) & ~excused
        assert not bad.any(), f"{mode} {what}: {int(bad.sum())}/{bad.size} rows over 2e-3 (max {err.max():.2e})"
        assert err.max() <= 5e-3, f"{mode} {what}: max {err.max():.2e}"
    else:
        cfg = TC_MODES[mode]
        assert np.isfinite(got).all()
        assert np.median(err) <= cfg["stress_median"], f"{mode} {what}: median {np.median(err):.2e} (regression guard)"
        assert err.max() <= cfg["stress_cap"], f"{mode} {what}: max {err.max():.2e} (regression guard)"


def _golden_case(path):
    g = np.load(path)
    spec = spec_of(g["meta"])
    n_steps, B, seed = (int(x) for x in g["meta"][5:8])
    gain = float(g["gain"])
    s, a = synth.states_actions(seed, B, spec.s_dim, spec.a_dim)
    probe = synth.rademacher(seed + 1, (B, spec.a_dim))
    probe_steps = synth.rademacher(seed + 2, (n_steps, B, spec.a_dim))
    return g, spec, n_steps, B, seed, gain, s, a, probe, probe_steps


def _caller(m, spec, st, at, role, n_steps, mode):
    if spec.n_heads == 2:
        return lambda **kw: m.log_prob(st, at, role, n_steps=n_steps, precision=mode, **kw)
    return lambda **kw: m.log_prob(torch.cat([st, at], dim=-1), n_steps=n_steps, precision=mode, **kw)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "logprob_*.npz"))), ids=os.path.basename)
def test_logprob_vs_reference_goldens(path):
    g, spec, n_steps, B, seed, gain, s, a, probe, probe_steps = _golden_case(path)
    st, at, pt, pst = dev(s), dev(a), dev(probe), dev(probe_steps)
    fp32_out = {}
    for mode, tol, kink in _modes():
        outs = {}
        for ri, role in enumerate(("expert", "agent")):
            flat = cf.make_params(spec, seed + (10 * ri if spec.n_heads == 1 else 0), gain)
            m = build_model(spec, flat)
            call = _caller(m, spec, st, at, role, n_steps, mode)
            got = call(probe=pt).cpu().numpy()
            assert_close(got, g[f"logp_{role}"], tol, f"{mode} {role} shared probe", kink)
            outs[role] = got
            got_s = call(probe=pst).cpu().numpy()
            assert_close(got_s, g[f"logp_{role}_stepprobe"], tol, f"{mode} {role} per-step probe", kink)
            got_x = None
            if f"logp_{role}_exact" in g:
                got_x = call(exact=True).cpu().numpy()
                assert_close(got_x, g[f"logp_{role}_exact"], 2 * tol, f"{mode} {role} exact trace", 2 * kink)
            if mode == "fp32":
                fp32_out[role] = (got, got_s, got_x)
            assert np.array_equal(got, call(probe=pt).cpu().numpy()), f"{mode} must be deterministic"
        if spec.n_heads == 2:  # fused two-role launch == two single-role launches, bit for bit
            le, la, rw = m.score(st, at, n_steps=n_steps, probe=pt, precision=mode)
            assert np.array_equal(le.cpu().numpy(), outs["expert"])
            assert np.array_equal(la.cpu().numpy(), outs["agent"])
            assert np.array_equal(rw.cpu().numpy(), outs["expert"] - outs["agent"])
    # ---- single-pass 16-bit tensor-core modes ----
    for mode in TC_MODES:
        for ri, role in enumerate(("expert", "agent")):
            flat = cf.make_params(spec, seed + (10 * ri if spec.n_heads == 1 else 0), gain)
            m = build_model(spec, flat)
            call = _caller(m, spec, st, at, role, n_steps, mode)
            f32 = fp32_out[role]
            _check_tc(call(probe=pt).cpu().numpy(), g[f"logp_{role}"], f32[0], mode, gain, f"{role} shared probe")
            _check_tc(call(probe=pst).cpu().numpy(), g[f"logp_{role}_stepprobe"], f32[1], mode, gain, f"{role} per-step probe")
            if f"logp_{role}_exact" in g:
                _check_tc(call(exact=True).cpu().numpy(), g[f"logp_{role}_exact"], f32[2], mode, gain, f"{role} exact trace")
            again = call(probe=pt).cpu().numpy()
            assert np.array_equal(again, call(probe=pt).cpu().numpy()), "tensor-core path must be deterministic"


def test_parity_modes_kink_rate_over_all_goldens():
    """The share of golden rows (all nets, roles, shared + per-step probes) outside the 1e-4 bar, per parity mode, next to
    the same figure for the float64 oracle (i.e. how often the REFERENCE's own fp32 arithmetic lands on a kink).  Bars are
    the measured rates with head-room (B200, round 2, 11520 rows: oracle 0.035 % = 4 rows, fp32 kernel 0.017 % = 2 rows,
    fp16x3 0.061 % = 7 rows; worst row 5.8e-4 / 8.8e-4 / 3.1e-4); every row stays under 2e-3."""
    from tests.gpu_common import rel_err
    over = {"oracle": 0, "fp32": 0, "fp16x3": 0}
    worst = dict(over)
    rows = 0
    for path in sorted(glob.glob(os.path.join(GOLDEN, "logprob_*.npz"))):
        g, spec, n_steps, B, seed, gain, s, a, probe, probe_steps = _golden_case(path)
        st, at, pt, pst = dev(s), dev(a), dev(probe), dev(probe_steps)
        for ri, role in enumerate(("expert", "agent")):
            flat = cf.make_params(spec, seed + (10 * ri if spec.n_heads == 1 else 0), gain)
            m = build_model(spec, flat)
            for key, pr_np, pr_dev in ((f"logp_{role}", probe, pt), (f"logp_{role}_stepprobe", probe_steps, pst)):
                rows += B
                errs = {"oracle": rel_err(cf.log_prob(spec, flat, s, a, role, pr_np, n_steps), g[key])}
                for mode in ("fp32", "fp16x3"):
                    errs[mode] = rel_err(_caller(m, spec, st, at, role, n_steps, mode)(probe=pr_dev).cpu().numpy(), g[key])
                for k, e in errs.items():
                    over[k] += int((e > 1e-4).sum())
                    worst[k] = max(worst[k], float(e.max()))
    rates = {k: v / rows for k, v in over.items()}
    print(f"kink rates over {rows} golden rows: {rates}; worst: {worst}")
    assert rates["fp32"] <= 1e-3 and rates["fp16x3"] <= 1.5e-3, (rates, worst)
    assert max(worst.values()) <= 2e-3, worst


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "vfield_*.npz"))), ids=os.path.basename)
def test_vfield_vs_reference_goldens(path):
    g = np.load(path)
    spec = spec_of(g["meta"])
    B, seed = int(g["meta"][6]), int(g["meta"][7])
    flat = cf.make_params(spec, seed, float(g["gain"]))
    s, a = synth.states_actions(seed, B, spec.s_dim, spec.a_dim)
    t = synth.rng(seed + 1).random((B, 1), dtype=np.float32)
    m = build_model(spec, flat)
    st, at, tt = dev(s), dev(a), dev(t)
    if spec.n_heads == 2:
        vc, r = m.net(at, st, tt)
        assert_close(vc.cpu().numpy(), g["v_c"], 1e-5, "v_c")
        assert_close(r.cpu().numpy(), g["r"], 1e-5, "tanh(r)")
        for role in ("expert", "agent"):
            assert_close(m.v_field(at, st, tt, role).cpu().numpy(), g[f"v_{role}"], 1e-5, role)
    else:
        assert_close(m.net(at, st, tt).cpu().numpy(), g["v"], 1e-5, "v")


@pytest.mark.parametrize("mode", ["fp16", "bf16"])
@pytest.mark.parametrize("B", [1, 127, 128, 129, 1000, 20000])
def test_tc_ragged_batches_smooth_net_meets_2e3(B, mode):
    """Tile edges of the tensor-core kernel (64-sample tiles, CTA pairs) at the 2e-3 bar -- every sample, no outlier
    allowance -- on a net at the reference's default-init scale."""
    spec = cf.NetSpec(17, 6, 256, 4, 2)
    flat = cf.make_params(spec, 78, 1.0)
    m = build_model(spec, flat)
    s, a = synth.states_actions(500 + B, B, 17, 6)
    probe = synth.rademacher(600 + B, (B, 6))
    le, la, rw = m.score(dev(s), dev(a), n_steps=16, probe=dev(probe), precision=mode)
    n = min(B, 512)
    for role, got in (("expert", le), ("agent", la)):
        ref = cf.log_prob(spec, flat, s[-n:], a[-n:], role, probe[-n:], 16)
        assert_close(got[-n:].cpu().numpy(), ref, 2e-3, f"{mode} B={B} {role}", 0.0, kink_cap=2e-3)   # every sample
    fe, fa, _ = m.score(dev(s), dev(a), n_steps=16, probe=dev(probe), precision="fp32")
    assert_close(le.cpu().numpy(), fe.cpu().numpy(), 2e-3, f"{mode} vs fp32 kernel", 0.0, kink_cap=2e-3)


@pytest.mark.parametrize("B", [1, 31, 32, 33, 127, 129, 1000])
def test_ragged_batches_vs_oracle(B):
    spec = cf.NetSpec(11, 3, 256, 4, 2)
    flat = cf.make_params(spec, 77, 2.0)
    m = build_model(spec, flat)
    s, a = synth.states_actions(100 + B, B, 11, 3)
    probe = synth.rademacher(200 + B, (B, 3))
    for mode, tol, kink in _modes():
        le, la, rw = m.score(dev(s), dev(a), n_steps=16, probe=dev(probe), precision=mode)
        for role, got in (("expert", le), ("agent", la)):
            ref = cf.log_prob(spec, flat, s, a, role, probe, 16)
            assert_close(got.cpu().numpy(), ref, tol, f"{mode} B={B} {role}", max(kink, 1.5 / B))


def test_empty_batch():
    spec = cf.NetSpec(6, 2, 128, 4, 2)
    m = build_model(spec, cf.make_params(spec, 5, 1.0))
    le, la, rw = m.score(torch.zeros(0, 6, device="cuda"), torch.zeros(0, 2, device="cuda"),
                         probe=torch.zeros(0, 2, device="cuda"))
    assert le.shape == (0,) and la.shape == (0,) and rw.shape == (0,)


@pytest.mark.parametrize("reward_type", ["raw", "norm", "exp", "clip", "smooth"])
def test_reward_types(reward_type):
    spec = cf.NetSpec(6, 2, 128, 4, 2)
    flat = cf.make_params(spec, 9, 3.0)
    m = build_model(spec, flat)
    B = 256
    s, a = synth.states_actions(31, B, 6, 2)
    probe = synth.rademacher(32, (B, 2))
    le, la, rw = m.score(dev(s), dev(a), n_steps=8, probe=dev(probe), reward_type=reward_type)
    r32 = (le - la).cpu().numpy()
    ref = cf.reward_transform(r32.astype(np.float32), reward_type)
    assert_close(rw.cpu().numpy(), ref, 2e-6, reward_type)


def test_bad_reward_type_raises():
    spec = cf.NetSpec(6, 2, 128, 4, 2)
    m = build_model(spec, cf.make_params(spec, 9, 1.0))
    with pytest.raises(ValueError):
        m.score(torch.zeros(4, 6, device="cuda"), torch.zeros(4, 2, device="cuda"), reward_type="bogus")


def test_default_probe_matches_reference_generator():
    """Without an explicit probe, scrf uses _rademacher((B,a)): identical on every call."""
    from modril_b200.flowril import _rademacher
    p1 = _rademacher((64, 3), device="cuda")
    p2 = _rademacher((64, 3), device="cuda")
    assert torch.equal(p1, p2) and set(p1.unique().tolist()) == {-1.0, 1.0}
    spec = cf.NetSpec(11, 3, 128, 4, 2)
    m = build_model(spec, cf.make_params(spec, 3, 2.0))
    s, a = synth.states_actions(1, 64, 11, 3)
    a1 = m.log_prob(dev(s), dev(a), "expert")
    a2 = m.log_prob(dev(s), dev(a), "expert", probe=p1)
    assert torch.equal(a1, a2)


def test_score_host_matches_device():
    import ctypes as C
    from modril_b200 import _native
    from modril_b200.flowril.pretrain import _t_mid
    spec = cf.NetSpec(17, 6, 256, 4, 2)
    flat = cf.make_params(spec, 21, 2.0)
    m = build_model(spec, flat)
    B, n = 3000, 12
    s, a = synth.states_actions(41, B, 17, 6)
    probe = synth.rademacher(42, (B, 6))
    le, la, rw = m.score(dev(s), dev(a), n_steps=n, probe=dev(probe), reward_type="smooth", precision="fp32")
    h = m.net.native()
    outs = [np.empty(B, np.float32) for _ in range(3)]
    ptr = lambda x: C.c_void_p(x.ctypes.data)
    rc = _native.lib().frl_score_host(h, h, ptr(s), ptr(a), ptr(probe), 0, _t_mid(n), n, 4, ptr(outs[0]),
                                      ptr(outs[1]), ptr(outs[2]), B, 0)
    _native.check(rc, "frl_score_host")
    assert np.array_equal(outs[0], le.cpu().numpy())
    assert np.array_equal(outs[1], la.cpu().numpy())
    assert np.array_equal(outs[2], rw.cpu().numpy())


@pytest.mark.parametrize("precision", ["fp32", "fp16", "fp16x3"])
@pytest.mark.parametrize("per_step", [False, True])
def test_score_host_pinned_and_pageable_buffers(precision, per_step):
    """Host entry point with pinned and with pageable buffers: bit-identical to the device-resident path, for shared
    and per-step probes and ragged sizes."""
    from modril_b200.flowril import score_pair_host
    spec = cf.NetSpec(11, 3, 256, 4, 2)
    m = build_model(spec, cf.make_params(spec, 22, 1.0))
    B, n = 5003, 8
    s, a = synth.states_actions(43, B, 11, 3)
    probe = synth.rademacher(44, (n, B, 3) if per_step else (B, 3))
    ref = m.score(dev(s), dev(a), n_steps=n, probe=dev(probe), precision=precision)
    hs, ha, hp = (torch.from_numpy(x).pin_memory() for x in (s, a, probe))
    outs = tuple(torch.empty(B).pin_memory() for _ in range(3))
    score_pair_host(m.net, m.net, hs, ha, hp, n_steps=n, precision=precision, out=outs)
    for got, want in zip(outs, ref):
        assert torch.equal(got, want.cpu())
    outs2 = tuple(torch.empty(B) for _ in range(3))
    score_pair_host(m.net, m.net, torch.from_numpy(s), torch.from_numpy(a), torch.from_numpy(probe), n_steps=n,
                    precision=precision, out=outs2)
    for got, want in zip(outs2, ref):
        assert torch.equal(got, want.cpu())


def test_full_size_properties():
    """BASELINE config 'maze, batch 64k, 20-step': properties that need no oracle at this size --
    (i) determinism, (ii) permutation equivariance over rows, (iii) agreement of a 4096-row slice with the oracle,
    (iv) exact trace == mean over many Hutchinson probes within its standard error."""
    spec = cf.NetSpec(6, 2, 256, 4, 2)
    flat = cf.make_params(spec, 50, 2.0)
    m = build_model(spec, flat)
    B, n = 65536, 20
    s, a = synth.states_actions(51, B, 6, 2)
    probe = synth.rademacher(52, (B, 2))
    st, at, pt = dev(s), dev(a), dev(probe)
    le, la, rw = m.score(st, at, n_steps=n, probe=pt)
    le2, la2, _ = m.score(st, at, n_steps=n, probe=pt)
    assert torch.equal(le, le2) and torch.equal(la, la2)
    perm = torch.randperm(B, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    lep, lap, _ = m.score(st[perm], at[perm], n_steps=n, probe=pt[perm])
    assert torch.equal(lep, le[perm]) and torch.equal(lap, la[perm])
    idx = slice(30000, 30000 + 1024)
    ref = cf.log_prob(spec, flat, s[idx], a[idx], "expert", probe[idx], n)
    assert_close(le[idx].cpu().numpy(), ref, 1e-4, "64k slice vs oracle", 0.011)
    assert torch.isfinite(rw).all()


@pytest.mark.parametrize("task,B,n_steps,mode", [("hopper", 262144, 32, "fp16"), ("halfcheetah", 262144, 32, "fp16"),
                                                 ("ant", 1048576, 32, "fp16"), ("hand", 1048576, 32, "fp16"),
                                                 ("hopper", 262144, 32, "fp16x3"), ("hand", 262144, 32, "fp16x3"),
                                                 ("ant", 1048576, 32, "fp16x3")])
def test_baseline_full_sizes_tensor_core_properties(task, B, n_steps, mode):
    """BASELINE.json configs at their full sizes on the tensor-core paths (fp16: one pass; fp16x3: the fp32-grade split):
    size-independent properties -- determinism, row-permutation equivariance (tiles regroup, results must not move),
    reward == logp_E - logp_A, and a slice agrees with the FP32 kernels / the float64 oracle within the mode's bar
    (2e-3 per sample for fp16 on these default-init-scale nets, 1e-4 for fp16x3)."""
    sd, ad = synth.TASK_DIMS[task]
    spec = cf.NetSpec(sd, ad, 256, 4, 2)
    flat = cf.make_params(spec, 60, 1.0)
    m = build_model(spec, flat)
    s, a = synth.states_actions(61, B, sd, ad)
    probe = synth.rademacher(62, (B, ad))
    st, at, pt = dev(s), dev(a), dev(probe)
    bar = 2e-3 if mode == "fp16" else 1e-4
    le, la, rw = m.score(st, at, n_steps=n_steps, probe=pt, precision=mode)
    le2, la2, rw2 = m.score(st, at, n_steps=n_steps, probe=pt, precision=mode)
    assert torch.equal(le, le2) and torch.equal(la, la2) and torch.equal(rw, rw2)
    assert torch.isfinite(le).all() and torch.isfinite(la).all()
    assert torch.equal(rw, le - la)
    perm = torch.randperm(B, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    lep, lap, _ = m.score(st[perm], at[perm], n_steps=n_steps, probe=pt[perm], precision=mode)
    assert torch.equal(lep, le[perm]) and torch.equal(lap, la[perm])
    lo = B // 3
    sl = slice(lo, lo + 2048)
    f32 = m.score(st[sl], at[sl], n_steps=n_steps, probe=pt[sl], precision="fp32")
    assert_close(le[sl].cpu().numpy(), f32[0].cpu().numpy(), bar, f"{task} {mode} vs fp32 kernels", kink_frac=0.0, kink_cap=bar)
    ref = cf.log_prob(spec, flat, s[lo:lo + 256], a[lo:lo + 256], "agent", probe[lo:lo + 256], n_steps)
    assert_close(la[lo:lo + 256].cpu().numpy(), ref, bar, f"{task} {mode} vs oracle", kink_frac=0.0, kink_cap=bar)
    assert_close(f32[1][:256].cpu().numpy(), ref, 1e-4, f"{task} fp32 vs oracle", kink_frac=0.0, kink_cap=1e-4)


@pytest.mark.parametrize("task,B,n_steps,mode", [("maze", 4097, 20, "fp16"), ("hopper", 20000, 32, "bf16"),
                                                 ("sine", 130, 8, "fp16"), ("pick", 9000, 32, "fp16"),
                                                 ("ant", 5000, 12, "fp16"), ("hand", 3000, 8, "bf16"),
                                                 ("ant111", 700, 5, "fp16")])
def test_cta_pair_integrator_is_bit_identical_to_single_cta_kernel(task, B, n_steps, mode, monkeypatch):
    """The default tensor-core kernel (CTA pairs, tcgen05 cta_group::2, csrc/logprob_tc3.cu) performs exactly the
    arithmetic of the 128-sample single-CTA kernel (FRL_TC_V1=1, csrc/logprob_tc.cu) -- same operand rounding, same
    bias folding, same K order -- so the two must agree bit for bit, also with per-step probes and the exact trace.
    (Any cross-CTA ordering bug would show up here as a mismatch.)  For a_dim >= 4 the divergence dot product is summed
    over the four threads of a sample as (d0+d1)+(d2+d3) instead of sequentially: 1-ulp differences, bound 2e-6."""
    sd, ad = synth.TASK_DIMS[task]
    spec = cf.NetSpec(sd, ad, 256, 4, 2)
    m = build_model(spec, cf.make_params(spec, 22, 1.0))
    s, a = synth.states_actions(33, B, sd, ad)
    probe = synth.rademacher(34, (B, ad))
    probe_steps = synth.rademacher(35, (n_steps, B, ad))
    calls = [dict(probe=dev(probe)), dict(probe=dev(probe_steps))] + ([dict(exact=True)] if ad <= 4 else [])
    for kw in calls:
        monkeypatch.setenv("FRL_TC_V1", "1")
        ref = m.score(dev(s), dev(a), n_steps=n_steps, precision=mode, **kw)
        monkeypatch.delenv("FRL_TC_V1")
        for _ in range(3):
            got = m.score(dev(s), dev(a), n_steps=n_steps, precision=mode, **kw)
            for g, r in zip(got[:2], ref[:2]):
                if ad <= 3:
                    assert torch.equal(g, r)
                else:
                    assert float(((g - r).abs() / r.abs().clamp_min(1.0)).max()) <= 2e-6


@pytest.mark.parametrize("task,B,n_steps", [("maze", 40000, 20), ("halfcheetah", 5000, 8)])
def test_private_and_shared_weight_streams_agree_bit_for_bit(task, B, n_steps, monkeypatch):
    """The CTA-pair kernel streams every layer's weight blocks once per tile slot (default) or once for both slots
    (FRL_TC3_SHARED=1); the two schedules reorder only the issue of independent MMAs, never the accumulation order."""
    sd, ad = synth.TASK_DIMS[task]
    spec = cf.NetSpec(sd, ad, 256, 4, 2)
    m = build_model(spec, cf.make_params(spec, 23, 1.0))
    s, a = synth.states_actions(36, B, sd, ad)
    probe = synth.rademacher(37, (B, ad))
    got = m.score(dev(s), dev(a), n_steps=n_steps, probe=dev(probe), precision="fp16")
    monkeypatch.setenv("FRL_TC3_SHARED", "1")
    ref = m.score(dev(s), dev(a), n_steps=n_steps, probe=dev(probe), precision="fp16")
    monkeypatch.delenv("FRL_TC3_SHARED")
    for g, r in zip(got, ref):
        assert torch.equal(g, r)


def test_score_host_wide_inputs_two_chunks():
    """Wide-input nets keep per-launch operand images in the handle; the host entry point alternates two streams over
    its 262144-row chunks, so the second chunk's images must wait for the first chunk's kernel (event in the handle)."""
    from modril_b200.flowril import score_pair_host
    spec = cf.NetSpec(132, 8, 256, 4, 2)
    m = build_model(spec, cf.make_params(spec, 24, 1.0))
    B, n = 262144 + 70001, 3
    s, a = synth.states_actions(45, B, 132, 8)
    probe = synth.rademacher(46, (B, 8))
    ref = m.score(dev(s), dev(a), n_steps=n, probe=dev(probe), precision="fp16")
    outs = score_pair_host(m.net, m.net, torch.from_numpy(s).pin_memory(), torch.from_numpy(a).pin_memory(),
                           torch.from_numpy(probe).pin_memory(), n_steps=n, precision="fp16")
    for got, want in zip(outs, ref):
        assert torch.equal(got, want.cpu())
