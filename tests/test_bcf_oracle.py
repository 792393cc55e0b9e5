"""CPU: the BCF-policy oracle (oracle/bcf_closed_form.py) against outputs of the reference's own VectorNetwork
(tests/golden/bcf_*.npz, written by oracle/gen_golden.py::gen_bcf)."""
import glob
import os

import numpy as np
import pytest

from oracle import bcf_closed_form as bcf
from oracle.gen_golden import bcf_inputs
from oracle import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "bcf_*.npz")))


def _case(path):
    g = np.load(path)
    s_dim, a_dim, hidden, tdim, B, n_steps, seed = (int(x) for x in g["meta"])
    spec = bcf.VNetSpec(s_dim, a_dim, hidden, tdim)
    return g, spec, bcf.make_params(spec, seed, 1.0), bcf_inputs(seed, B, s_dim, a_dim), n_steps


def test_fixtures_present():
    assert len(CASES) >= 5 and len(glob.glob(os.path.join(GOLDEN, "bcftrain_*.npz"))) >= 2


@pytest.mark.parametrize("path", CASES, ids=os.path.basename)
def test_oracle_matches_reference_vector_network(path):
    g, spec, flat, (s, a, x0, t), n_steps = _case(path)
    coef = float(g["coef"])
    v = bcf.forward(spec, flat.astype(np.float64), x0.astype(np.float64), s.astype(np.float64), t.astype(np.float64).reshape(-1, 1))
    assert np.abs(v - g["v_f64"]).max() <= 1e-12
    assert np.abs(v - g["v_f32"]).max() <= 2e-5
    total, flow, act, grad = bcf.bcf_loss_and_grad(spec, flat, s, a, x0, t, coef)
    assert np.allclose([total, flow, act], g["loss_f64"], rtol=1e-12)
    assert np.allclose([total, flow, act], g["loss_f32"], rtol=2e-6)
    ref = g["grad_f64"]
    tol = 1e-12 if ref.dtype == np.float64 else 2e-7
    assert np.abs(grad - ref).max() <= tol * max(1.0, np.abs(ref).max())
    ref32 = g["grad_f32"].astype(np.float64)
    assert np.sqrt(((grad - ref32) ** 2).sum()) <= 1e-4 * np.sqrt((ref32 ** 2).sum())
    x = bcf.sample_actions(spec, flat, s, x0, n_steps)
    assert np.abs(x - g["sample_f64"]).max() <= 1e-11
    assert np.abs(x - g["sample_f32"]).max() <= 5e-5


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "bcftrain_00*.npz"))), ids=os.path.basename)
def test_oracle_bc_updates_follow_reference_curve(path):
    """First 40 updates of _bc_step + _standard_step (float64 oracle vs the reference's fp32 run)."""
    g = np.load(path)
    s_dim, a_dim, hidden, tdim, steps, B, seed = (int(x) for x in g["meta"])
    spec = bcf.VNetSpec(s_dim, a_dim, hidden, tdim)
    flat = bcf.make_params(spec, seed, 1.0).astype(np.float64)
    m, v = np.zeros_like(flat), np.zeros_like(flat)
    pool_s, _ = synth.states_actions(seed + 1, 4096, s_dim, a_dim)
    pool_a = synth.expert_actions(seed + 2, pool_s, a_dim)
    for it in range(40):
        r = synth.rng(seed + 100 + it)
        i = r.integers(0, 4096, B)
        x0 = r.standard_normal((B, a_dim)).astype(np.float32)
        t = r.random(B).astype(np.float32)
        flat, m, v, losses, norm = bcf.bc_step(spec, flat, m, v, it + 1, pool_s[i], pool_a[i], x0, t, float(g["lr"]))
        # Adam with eps 1e-12 amplifies fp32-vs-fp64 rounding step by step: tight while the runs coincide, loose later
        assert np.allclose(losses, g["curve"][it], rtol=1e-4 if it < 10 else 3e-3), (it, losses, g["curve"][it])
        assert abs(norm - g["gradnorm"][it]) <= (2e-3 if it < 10 else 2e-2) * g["gradnorm"][it]
