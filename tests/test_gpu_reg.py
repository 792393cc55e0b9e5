"""GPU parity of the anti-divergence / Jacobian-stability regularisers through the C ABI:
`frl_reg_loss_grad`, `frl_grad_axpy`, `frl_tensor_norms`, and the regularised update loop of flowril.py:251-330."""
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


def _grad_close(grad, ref, what, l2_tol=1e-4, max_tol=1e-3):
    ref = np.asarray(ref, np.float64)
    d = np.asarray(grad, np.float64) - ref
    l2 = np.sqrt((d ** 2).sum()) / np.sqrt((ref ** 2).sum())
    mx = np.abs(d).max() / np.abs(ref).max()
    assert l2 <= l2_tol and mx <= max_tol, f"{what}: rel L2 {l2:.3e}, max {mx:.3e}"


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "reg_*.npz"))), ids=os.path.basename)
def test_reg_losses_and_grads_vs_reference(path):
    from modril_b200.flowril import tensor_grad_norm_sum
    g = np.load(path)
    spec = spec_of(g["meta"])
    N, seed = int(g["meta"][6]), int(g["meta"][7])
    flat = cf.make_params(spec, seed, float(g["gain"]))
    s, a = synth.states_actions(seed, N, spec.s_dim, spec.a_dim)
    t = synth.rng(seed + 1).random((N, 1), dtype=np.float32)
    probe = synth.rademacher(seed + 2, (N, spec.a_dim))
    m = build_model(spec, flat)
    out = m.reg_losses(dev(s), dev(a), t=dev(t), probe=dev(probe), need_grad=True)
    assert_close(np.array(out["loss_anti"].item()), g["loss_anti"], 2e-5, "loss_anti")
    assert_close(np.array(out["loss_stable"].item()), g["loss_stable"], 2e-5, "loss_stable")
    ola, ols, oga, ogs = cf.reg_losses_and_grads(spec, flat, s, a, t, probe)
    for name, ref64 in (("anti", oga), ("stable", ogs)):
        grad = out[f"grad_{name}"].cpu().numpy()
        norms, off = [], 0
        for _, shape in spec.shapes():
            n = int(np.prod(shape))
            norms.append(np.sqrt((grad[off:off + n].astype(np.float64) ** 2).sum()))
            off += n
        assert np.allclose(norms, g[f"tensor_gradnorms_{name}"], rtol=3e-4, atol=1e-6), (name, norms)
        ns = tensor_grad_norm_sum(m.net, out[f"grad_{name}"]).item()     # GradNorm2D's quantity (networks.py:111)
        assert abs(ns - g[f"normsum_{name}"]) <= 3e-4 * g[f"normsum_{name}"]
        if f"grad_{name}" in g:
            ref = g[f"grad_{name}"]
            assert np.abs(grad - ref).max() <= 2e-4 * np.abs(ref).max(), np.abs(grad - ref).max() / np.abs(ref).max()
        _grad_close(grad, ref64, f"{name} vs float64 oracle")


@pytest.mark.parametrize("N", [1, 5, 64, 130, 1000, 4097])
def test_reg_ragged_vs_oracle(N):
    spec = cf.NetSpec(17, 6, 256, 4, 2)
    flat = cf.make_params(spec, 92, 2.0)
    m = build_model(spec, flat)
    s, a = synth.states_actions(500 + N, N, 17, 6)
    t = synth.rng(600 + N).random((N, 1), dtype=np.float32)
    probe = synth.rademacher(700 + N, (N, 6))
    out = m.reg_losses(dev(s), dev(a), t=dev(t), probe=dev(probe), need_grad=True)
    la, ls, ga, gs = cf.reg_losses_and_grads(spec, flat, s, a, t, probe)
    assert abs(out["loss_anti"].item() - la) <= 2e-5 * max(1.0, abs(la))
    assert abs(out["loss_stable"].item() - ls) <= 2e-5 * max(1.0, abs(ls))
    _grad_close(out["grad_anti"].cpu().numpy(), ga, f"anti N={N}")
    _grad_close(out["grad_stable"].cpu().numpy(), gs, f"stable N={N}")


def test_reg_single_terms_deterministic_and_depth():
    """Each term alone equals the term of the joint call; repeated calls are bit-identical; other depths / widths."""
    for (hidden, depth) in ((128, 2), (256, 3), (128, 1)):
        spec = cf.NetSpec(11, 3, hidden, depth, 2)
        flat = cf.make_params(spec, 17, 2.0)
        m = build_model(spec, flat)
        N = 300
        s, a = synth.states_actions(9, N, 11, 3)
        t = synth.rng(10).random((N, 1), dtype=np.float32)
        probe = synth.rademacher(11, (N, 3))
        args = (dev(s), dev(a))
        kw = dict(t=dev(t), probe=dev(probe), need_grad=True)
        both = {k: v.clone() for k, v in m.reg_losses(*args, **kw).items()}
        again = m.reg_losses(*args, **kw)
        for k in both:
            assert torch.equal(both[k], again[k]), k
        only_a = {k: v.clone() for k, v in m.reg_losses(*args, stable=False, **kw).items()}
        only_s = {k: v.clone() for k, v in m.reg_losses(*args, anti=False, **kw).items()}
        assert set(only_a) == {"loss_anti", "grad_anti"} and set(only_s) == {"loss_stable", "grad_stable"}
        assert torch.equal(only_a["grad_anti"], both["grad_anti"]) and torch.equal(only_s["grad_stable"], both["grad_stable"])
        la, ls, ga, gs = cf.reg_losses_and_grads(spec, flat, s, a, t, probe)
        _grad_close(both["grad_anti"].cpu().numpy(), ga, f"anti d{depth}")
        _grad_close(both["grad_stable"].cpu().numpy(), gs, f"stable d{depth}")


def test_reg_rejects_single_head_net():
    from modril_b200.flowril.pretrain import reg_losses_native
    spec = cf.NetSpec(6, 2, 128, 4, 1)
    m = build_model(spec, cf.make_params(spec, 1, 1.0))
    z = torch.zeros(4, 2, device="cuda")
    with pytest.raises(ValueError):
        reg_losses_native(m.net, torch.zeros(4, 6, device="cuda"), z, torch.zeros(4, 1, device="cuda"), z)


def test_grad_axpy_gate_and_coef():
    from modril_b200.flowril.pretrain import _axpy
    y = torch.ones(1000, device="cuda")
    x = torch.full((1000,), 2.0, device="cuda")
    num, den = torch.tensor([3.0], device="cuda"), torch.tensor([0.5], device="cuda")
    gate = torch.tensor([1e-4, 0.0], device="cuda")
    _axpy(y, x, 1e-3, num=num, den=den, eps=1e-8, gate=gate, coef_out=gate)
    c = 1e-3 * 3.0 / (0.5 + 1e-8)
    assert abs(gate[0].item() - c) < 1e-9 and gate[1].item() == 0.0
    assert torch.allclose(y, torch.full_like(y, 1.0 + 2.0 * c))
    gate[0] = 0.0                                  # a weight that hit zero stays off (use_anti = w > 0)
    _axpy(y, x, 1e-3, num=num, den=den, eps=1e-8, gate=gate, coef_out=gate)
    assert gate[0].item() == 0.0 and torch.allclose(y, torch.full_like(y, 1.0 + 2.0 * c))
    _axpy(y, x, 0.5)                               # plain fixed weight
    assert torch.allclose(y, torch.full_like(y, 2.0 + 2.0 * c))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "trainreg_*.npz"))), ids=os.path.basename)
def test_regularised_training_curves_vs_reference(path):
    """flowril.py:251-284 + :311-330 with the regularisers (warm-up weighting / fixed weights) on the native route.
    Tolerances as in test_gpu_train: per-step values over the first 50 steps, 50-step moving averages to the end."""
    from modril_b200.flowril import FusedAdam, train_step
    g = np.load(path)
    spec = spec_of(g["meta"])
    steps, B, seed = (int(x) for x in g["meta"][5:8])
    warm = int(g["mode"]) == 0
    m = build_model(spec, cf.make_params(spec, seed, 1.0))
    opt = FusedAdam([m.net], lr=float(g["lr"]))
    pool_s, _ = synth.states_actions(seed + 1, 4096, spec.s_dim, spec.a_dim)
    pool_aE = synth.expert_actions(seed + 2, pool_s, spec.a_dim)
    pool_aA = np.clip(synth.rng(seed + 3).standard_normal((4096, spec.a_dim)), -1, 1).astype(np.float32)
    d_s, d_aE, d_aA = dev(pool_s), dev(pool_aE), dev(pool_aA)
    probe = dev(synth.rademacher(seed + 4, (2 * B, spec.a_dim)))
    gate_a = torch.tensor([1e-4, 0.0], device="cuda")
    gate_s = torch.tensor([1e-6, 0.0], device="cuda")
    log, wlog = [], []
    for it in range(steps):
        r = synth.rng(seed + 100 + it)
        iE, iA = dev(r.integers(0, 4096, B)), dev(r.integers(0, 4096, B))
        tE, nE = synth.cfm_draws(seed + 10000 + 2 * it, B, spec.a_dim)
        tA, nA = synth.cfm_draws(seed + 10001 + 2 * it, B, spec.a_dim)
        t_mix = synth.rng(seed + 70000 + it).random((2 * B, 1), dtype=np.float32)
        terms = [dict(net=m.net, sign=1.0, s=d_s[iE], a0=d_aE[iE], t=dev(tE), noise=dev(nE), rate=1.0),
                 dict(net=m.net, sign=-1.0, s=d_s[iA], a0=d_aA[iA], t=dev(tA), noise=dev(nA), rate=1.0)]
        spec_w = ("warmup", 1e-3, 1e-8)
        reg = dict(net=m.net, s=torch.cat([d_s[iE], d_s[iA]]), a=torch.cat([d_aE[iE], d_aA[iA]]), t=dev(t_mix),
                   probe=probe, anti=spec_w if warm else 1e-4, stable=spec_w if warm else 1e-6,
                   gate_anti=gate_a, gate_stable=gate_s)
        lE, lA = train_step(terms, opt, max_norm=1.0, reg=reg)
        log.append(torch.stack([lE, lA, reg["out"]["loss_anti"], reg["out"]["loss_stable"]]))
        wlog.append(torch.stack([gate_a[0], gate_s[0]]).clone())
    curve = torch.stack(log).cpu().numpy()
    weights = torch.stack(wlog).cpu().numpy()
    ref, refw = g["curve"], g["weights"]
    head = min(50, steps)
    assert_close(curve[:head], ref[:head], 2e-3, "first steps")
    relw = np.abs(weights[:head] - refw[:head]) / np.abs(refw[:head])
    assert relw.max() <= 5e-3, relw.max()
    k = min(50, steps)
    ma = lambda x: np.stack([np.convolve(x[:, j], np.ones(k) / k, mode="valid") for j in range(x.shape[1])], 1)
    rel = np.abs(ma(curve) - ma(ref)) / np.abs(ma(ref))
    assert rel[:, :2].max() <= 3e-2 and rel[:, 2:].max() <= 0.1, rel.max(axis=0)
