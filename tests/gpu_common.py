"""Shared helpers for the -m gpu parity tests (all compute goes through the C ABI via modril_b200)."""
import numpy as np
import torch

from oracle import closed_form as cf


def rel_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return np.abs(got - ref) / np.maximum(1.0, np.abs(ref))


def assert_close(got, ref, rel, what, kink_frac=0.0, kink_cap=5e-3):
    """|got-ref| <= rel*max(1,|ref|) for all but `kink_frac` of the entries, and <= kink_cap for all.
    (ReLU-kink outliers: see tests/test_oracle_vs_golden.py::_close and DESIGN.md.)"""
    err = rel_err(got, ref)
    bad = err > rel
    assert bad.mean() <= kink_frac, f"{what}: {int(bad.sum())}/{bad.size} over {rel:g}; max {err.max():.3e}"
    assert err.max() <= kink_cap, f"{what}: max err {err.max():.3e} over the kink cap {kink_cap:g}"


def spec_of(meta):
    s_dim, a_dim, hidden, depth, n_heads = (int(x) for x in meta[:5])
    return cf.NetSpec(s_dim, a_dim, hidden, depth, n_heads)


def build_model(spec, flat, device="cuda:0"):
    """modril_b200 estimator with the given flat weights (state_dict order)."""
    from modril_b200.flowril import CoupledResidualFM, FlowMatching

    if spec.n_heads == 2:
        m = CoupledResidualFM(spec.s_dim, spec.a_dim, spec.hidden, depth=spec.depth)
    else:
        m = FlowMatching(spec.s_dim, spec.a_dim, spec.hidden, depth=spec.depth)
    m = m.to(device)
    sd = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in cf.unpack(spec, np.asarray(flat, np.float32)).items()}
    m.load_state_dict(sd)
    return m


def dev(x, device="cuda:0"):
    return torch.from_numpy(np.ascontiguousarray(x)).to(device)
