"""CPU cross-checks between the two restatements of the reference arithmetic that do not need the reference checkout:
``oracle/torch_port.py`` (op-for-op torch port: cat -> Linear / ReLU -> heads -> ONE autograd VJP per Euler step; the
fallback CPU baseline of bench.py) against ``oracle/closed_form.py`` (numpy float64, forward-mode tangent, hand-written
backward), and -- when ``oracle/_ref`` is staged -- the reference's own unmodified classes against both."""
import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import synth
from oracle import torch_port as tp


@pytest.mark.parametrize("dims", [(11, 3, 128, 4, 2), (6, 2, 256, 3, 2)])
def test_torch_port_log_prob_and_loss_match_the_numpy_oracle(dims):
    spec = cf.NetSpec(*dims)
    flat = cf.make_params(spec, 3, 2.0)
    B, n = 48, 12
    s, a = synth.states_actions(4, B, spec.s_dim, spec.a_dim)
    ft, st, at = torch.from_numpy(flat), torch.from_numpy(s), torch.from_numpy(a)
    for probe in (synth.rademacher(5, (B, spec.a_dim)), synth.rademacher(6, (n, B, spec.a_dim))):
        for role in ("expert", "agent"):
            got = tp.log_prob(spec, ft, st, at, role, torch.from_numpy(probe), n).detach().numpy().astype(np.float64)
            want = cf.log_prob(spec, flat, s, a, role, probe, n)
            err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
            assert np.sort(err)[-2] <= 2e-5 and err.max() <= 1e-3, err.max()   # one ReLU-kink row allowed
    t, noise = synth.cfm_draws(7, B, spec.a_dim)
    flat_t = torch.from_numpy(flat).clone().requires_grad_(True)
    loss = tp.fm_loss(spec, flat_t, st, at, "agent", torch.from_numpy(t), torch.from_numpy(noise))
    loss.backward()
    want_loss, want_grad = cf.fm_loss_and_grad(spec, flat, s, a, "agent", t, noise)
    assert abs(loss.item() - want_loss) <= 2e-5 * max(1.0, abs(want_loss))
    g = flat_t.grad.numpy().astype(np.float64)
    assert np.sqrt(((g - want_grad) ** 2).sum()) <= 1e-4 * np.sqrt((want_grad ** 2).sum())


def test_staged_reference_classes_match_the_oracle():
    """The unmodified reference modules staged under oracle/_ref (what bench.py --impl reference times) compute what the
    oracle computes, on the probe the reference itself draws (its shape-deterministic ``_rademacher``)."""
    from oracle import build_ref

    P = build_ref.load()
    if P is None:
        pytest.skip("oracle/_ref is not staged (python -m oracle.build_ref in the build container)")
    spec = cf.NetSpec(11, 3, 128, 4, 2)
    flat = cf.make_params(spec, 9, 1.0)
    model = P.CoupledResidualFM(11, 3, 128, depth=4)
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in cf.unpack(spec, flat).items()})
    s, a = synth.states_actions(10, 40, 11, 3)
    st, at = torch.from_numpy(s), torch.from_numpy(a)
    probe = P._rademacher((40, 3), device=st.device).numpy()
    for role in ("expert", "agent"):
        with torch.enable_grad():
            got = model.log_prob(st, at, role, n_steps=8).detach().numpy().astype(np.float64)
        want = cf.log_prob(spec, flat, s, a, role, probe, 8)
        err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
        assert np.sort(err)[-2] <= 2e-5 and err.max() <= 1e-3, err.max()
