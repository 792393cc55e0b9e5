"""Free ``--hidden-dim`` (reference flowril.py:562; the pre-trainer's CLI default is 128, its sweeps use others): trunk
widths that are not one of the kernels' native 128 / 256 run zero-padded to the next of the two
(modril_b200/flowril/networks.py::_pad_plan).  Checked against the float64 oracle evaluated at the TRUE width: velocity
field, log-density (fp32 and the fp16x3 default), CFM loss + full gradient through autograd, the regularisers, and
whole updates through ``train_step`` (fp32 and tensor-core) against the oracle's clip + Adam."""
import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import synth
from tests.gpu_common import assert_close, build_model, dev

pytestmark = pytest.mark.gpu


def _rel_l2(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return np.sqrt(((got - ref) ** 2).sum()) / np.sqrt((ref ** 2).sum())


@pytest.mark.parametrize("hidden,depth,heads", [(64, 4, 2), (100, 3, 2), (192, 4, 2), (200, 2, 1), (7, 2, 2)])
def test_logprob_loss_and_gradients_at_free_hidden_dims(hidden, depth, heads):
    spec = cf.NetSpec(11, 3, hidden, depth, heads)
    flat = cf.make_params(spec, 40 + hidden, 1.0)
    m = build_model(spec, flat)
    assert m.net.native_hidden() in (128, 256) and m.net.flat_params().numel() == flat.size
    B, n = 333, 8
    s, a = synth.states_actions(3, B, 11, 3)
    probe = synth.rademacher(4, (B, 3))
    role = "expert"
    for prec in ("fp32", "fp16x3"):
        if heads == 2:
            got = m.log_prob(dev(s), dev(a), role, n_steps=n, probe=dev(probe), precision=prec)
        else:
            got = m.log_prob(torch.cat([dev(s), dev(a)], dim=-1), n_steps=n, probe=dev(probe), precision=prec)
        ref = cf.log_prob(spec, flat, s, a, role, probe, n)
        assert_close(got.cpu().numpy(), ref, 1e-4, f"log_prob {prec} H={hidden}")
    t, noise = synth.cfm_draws(5, B, 3)
    if heads == 2:
        loss = m.fm_loss(dev(s), dev(a), role, t=dev(t), noise=dev(noise))
        ref_loss, ref_grad = cf.fm_loss_and_grad(spec, flat, s, a, role, t, noise)
    else:
        dq = synth.rng(6).standard_normal((B, 3)).astype(np.float32)
        loss = m.fm_loss(dev(s), dev(a), dequant=dev(dq), t=dev(t), noise=dev(noise))
        ref_loss, ref_grad = cf.fm_loss_and_grad(spec, flat, s, a, role, t, noise, dequant=dq * np.float32(0.02))
    loss.backward()
    grad = torch.cat([p.grad.reshape(-1) for p, _, _ in m.net.param_slices()]).cpu().numpy()
    assert abs(float(loss) - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss))
    assert _rel_l2(grad, ref_grad) <= 1e-4


@pytest.mark.parametrize("prec,tol", [("fp32", 2e-5), ("fp16", 2e-2)])
def test_updates_at_free_hidden_dim_follow_the_oracle(prec, tol):
    from modril_b200.flowril import FusedAdam, train_step

    spec = cf.NetSpec(17, 6, 192, 4, 2)
    flat = cf.make_params(spec, 9, 1.0)
    m = build_model(spec, flat)
    opt = FusedAdam([m.net], lr=1e-3)
    B = 512
    s, a = synth.states_actions(11, B, 17, 6)
    draws = [synth.cfm_draws(12 + i, B, 6) for i in range(2)]
    ref, mom, var = flat.astype(np.float64).copy(), np.zeros(flat.size), np.zeros(flat.size)
    for step in range(1, 4):
        losses = train_step([dict(net=m.net, sign=sg, s=dev(s), a0=dev(a), t=dev(t), noise=dev(z), rate=1.0)
                             for sg, (t, z) in zip((1.0, -1.0), draws)], opt, 1.0, precision=prec)
        ref, mom, var, ref_losses, _ = cf.train_step(spec, ref, mom, var, step,
                                                     [("expert", s, a, *draws[0]), ("agent", s, a, *draws[1])], 1e-3, 1.0)
        for got, want in zip(losses, ref_losses):
            assert abs(float(got) - want) <= max(tol, 1e-4) * max(1.0, abs(want))
    assert _rel_l2(m.net.flat_params().cpu().numpy(), ref) <= tol
    # the regularisers at the true width
    t_mix = synth.rng(20).random((B, 1)).astype(np.float32)
    probe = synth.rademacher(21, (B, 6))
    flat_now = m.net.flat_params().detach().cpu().numpy().copy()
    out = m.reg_losses(dev(s), dev(a), t=dev(t_mix), probe=dev(probe), need_grad=True)
    la, ls, ga, gs = cf.reg_losses_and_grads(spec, flat_now, s, a, t_mix, probe)
    assert abs(float(out["loss_anti"]) - la) <= 1e-3 * abs(la) + 1e-9
    assert abs(float(out["loss_stable"]) - ls) <= 1e-3 * abs(ls) + 1e-9
    assert _rel_l2(out["grad_anti"].cpu().numpy(), ga) <= 1e-3 and _rel_l2(out["grad_stable"].cpu().numpy(), gs) <= 1e-3
    assert out["grad_anti"].numel() == flat.size
