"""Host-side logic that needs no GPU: drop-in surface (state_dict keys, flat views, freezing), oracle helpers vs torch,
and the torch-CPU port (bench baseline) vs the numpy oracle."""
import copy

import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import synth, torch_port


def test_state_dict_keys_match_reference_layout():
    from modril_b200.flowril import CoupledResidualFM, FlowMatching
    m = CoupledResidualFM(11, 3, 256)
    spec = cf.NetSpec(11, 3, 256, 4, 2)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == spec.shapes()
    m2 = FlowMatching(6, 2, 128, depth=3)
    spec2 = cf.NetSpec(6, 2, 128, 3, 1)
    assert [(k, tuple(v.shape)) for k, v in m2.state_dict().items()] == spec2.shapes()
    assert m.net.flat_params().numel() == spec.n_params()


def test_parameters_are_views_of_one_flat_buffer_and_survive_load_and_copy():
    from modril_b200.flowril import CoupledResidualFM
    m = CoupledResidualFM(6, 2, 128)
    flat = m.net.flat_params()
    off = 0
    for p in m.parameters():
        assert p.data_ptr() == flat.data_ptr() + 4 * off
        off += p.numel()
    spec = cf.NetSpec(6, 2, 128, 4, 2)
    new = cf.make_params(spec, 3, 1.0)
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in cf.unpack(spec, new).items()})
    assert np.array_equal(m.net.flat_params().numpy(), new)
    m2 = copy.deepcopy(m)
    assert np.array_equal(m2.net.flat_params().numpy(), new)
    assert m2.net.flat_params().data_ptr() != m.net.flat_params().data_ptr()
    with torch.no_grad():
        m.net.head_r.weight.mul_(2.0)   # in-place edits through the Parameter are edits of the flat buffer
    assert np.allclose(m.net.flat_params().numpy()[-(2 * 128 + 2):-2], 2 * new[-(2 * 128 + 2):-2])


def test_cpu_tensors_raise_no_fallback():
    from modril_b200.flowril import CoupledResidualFM
    m = CoupledResidualFM(6, 2, 128)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.log_prob(torch.zeros(4, 6), torch.zeros(4, 2), "expert")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.fm_loss(torch.zeros(4, 6), torch.zeros(4, 2), "expert")


def test_fused_adam_state_dict_layout_and_freezing():
    from modril_b200.flowril import CoupledResidualFM, FusedAdam
    m = CoupledResidualFM(6, 2, 128)
    for p in m.net.shared.parameters():      # reference flowril.py:52-58
        p.requires_grad = False
    for p in m.net.head_c.parameters():
        p.requires_grad = False
    opt = FusedAdam([m.net], lr=1e-3)
    sd = opt.state_dict()
    assert len(sd["param_groups"][0]["params"]) == 2          # head_r.weight, head_r.bias
    ref = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-3)
    assert set(sd["param_groups"][0].keys()) >= {"lr", "betas", "eps", "weight_decay", "amsgrad"}
    assert sd["param_groups"][0]["betas"] == ref.state_dict()["param_groups"][0]["betas"]
    arr, n = opt._segments([m.net.flat_grad()])
    assert n == 1 and arr[0].n == 2 * 128 + 2                  # one contiguous trainable run at the tail


def test_oracle_linspace_matches_torch():
    for n in (1, 2, 7, 20, 32, 33, 100, 256):
        g = torch.linspace(0.0, 1.0, n + 1)
        assert np.array_equal(cf.torch_linspace_f32(n), g.numpy()), n
        tm = (0.5 * (g[:-1] + g[1:])).numpy()
        assert np.array_equal(cf.t_mid_grid(n), tm), n


def test_t_mid_binding_matches_oracle():
    from modril_b200.flowril.pretrain import _t_mid
    for n in (20, 32):
        assert np.array_equal(np.array(list(_t_mid(n, 'cpu')), np.float32), cf.t_mid_grid(n))


def test_reward_transform_matches_torch_formulas():
    r = torch.linspace(-30, 30, 121)
    s = torch.sigmoid(r)
    ref = {"raw": r, "norm": (s + 1e-20).log() - (1 - s + 1e-20).log(), "exp": -torch.exp(r) / (1 + torch.exp(r)),
           "clip": torch.clamp(r, -20.0, 20.0), "smooth": -torch.log1p(torch.exp(-r))}   # reference flowril.py:195-205
    for k, v in ref.items():
        got = cf.reward_transform(r.numpy(), k)
        assert np.allclose(got, v.numpy(), rtol=2e-6, atol=1e-6, equal_nan=True), k
    with pytest.raises(ValueError):
        cf.reward_transform(r.numpy(), "bogus")


def test_torch_port_matches_numpy_oracle():
    spec = cf.NetSpec(11, 3, 128, 4, 2)
    flat = cf.make_params(spec, 4, 2.0)
    B, n = 64, 8
    s, a = synth.states_actions(1, B, 11, 3)
    probe = synth.rademacher(2, (B, 3))
    ft = torch.from_numpy(flat)
    le, la, rw = torch_port.score(spec, ft, ft, torch.from_numpy(s), torch.from_numpy(a), torch.from_numpy(probe), n)
    for role, got in (("expert", le), ("agent", la)):
        ref = cf.log_prob(spec, flat, s, a, role, probe, n)
        assert np.abs(got.detach().numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
    t, noise = synth.cfm_draws(3, B, 3)
    loss = torch_port.fm_loss(spec, ft.clone().requires_grad_(True), torch.from_numpy(s), torch.from_numpy(a), "agent",
                              torch.from_numpy(t), torch.from_numpy(noise))
    ref_loss, _ = cf.fm_loss_and_grad(spec, flat, s, a, "agent", t, noise)
    assert abs(float(loss) - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss))


def test_oracle_training_curve_vs_reference_golden(golden_dir):
    """The numpy oracle's hand-written backward + clip + Adam reproduces the reference's 50-step fine-tune run."""
    import os
    g = np.load(os.path.join(golden_dir, "train_01_hopper_scrf_ft_h128_50steps.npz"))
    s_dim, a_dim, hidden, depth, n_heads, steps, B, seed = (int(x) for x in g["meta"])
    spec = cf.NetSpec(s_dim, a_dim, hidden, depth, n_heads)
    flat = cf.make_params(spec, seed, 1.0).astype(np.float64)
    trainable = np.zeros(flat.size, bool)
    n_frozen = sum(int(np.prod(sh)) for name, sh in spec.shapes() if "head_r" not in name)
    trainable[n_frozen:] = True
    m, v = np.zeros_like(flat), np.zeros_like(flat)
    pool_s, _ = synth.states_actions(seed + 1, 4096, s_dim, a_dim)
    pool_aE = synth.expert_actions(seed + 2, pool_s, a_dim)
    pool_aA = np.clip(synth.rng(seed + 3).standard_normal((4096, a_dim)), -1, 1).astype(np.float32)
    for it in range(20):
        r = synth.rng(seed + 100 + it)
        iE, iA = r.integers(0, 4096, B), r.integers(0, 4096, B)
        tE, nE = synth.cfm_draws(seed + 10000 + 2 * it, B, a_dim)
        tA, nA = synth.cfm_draws(seed + 10001 + 2 * it, B, a_dim)
        flat, m, v, losses, norm = cf.train_step(
            spec, flat, m, v, it + 1,
            [("expert", pool_s[iE], pool_aE[iE], tE, nE), ("agent", pool_s[iA], pool_aA[iA], tA, nA)],
            lr=float(g["lr"]), max_norm=1.0, trainable=trainable)
        assert np.allclose(losses, g["curve"][it], rtol=2e-4), (it, losses, g["curve"][it])


def test_gradnorm2d_mirror_follows_the_reference_trajectory(golden_dir):
    """modril_b200's GradNorm2D (host-side scalar logic) against a trajectory of the reference's class
    (networks.py:52-138; fixture from oracle/gen_golden.py gen_gradnorm).  The reference obtains the two gradient-norm
    sums with autograd; here they are passed in (closed form for the toy losses of the fixture)."""
    import os

    import torch

    from modril_b200.flowril.networks import GradNorm2D

    g = np.load(os.path.join(golden_dir, "gradnorm_00.npz"))
    gn = GradNorm2D(beta_init=1e-4, gamma_init=1e-6, alpha=0.1, warmup_steps=20, update_freq=5, ema_decay=0.9,
                    beta_min=1e-8, beta_max=1e1, gamma_min=1e-12, gamma_max=1e-2)
    p1, p2 = np.array([1.0, -2.0, 0.5]), np.array([0.3, 0.7])
    n_anti = np.linalg.norm(2 * p1) + np.linalg.norm(2 * p2)          # sum of per-tensor norms (networks.py:111)
    n_jpen = np.linalg.norm(4 * p1 ** 3) + np.linalg.norm(np.sign(p2))
    va, vj = (p1 ** 2).sum() + (p2 ** 2).sum(), (p1 ** 4).sum() + np.abs(p2).sum()
    for k in range(len(g["ca"])):
        ca, cb = float(g["ca"][k]), float(g["cb"][k])
        la, lj = torch.tensor(ca * va, dtype=torch.float32), torch.tensor(cb * vj, dtype=torch.float32)
        reg, beta, gamma = gn(la, lj)
        gn.update(la, lj, None, g_anti=ca * n_anti, g_jpen=cb * n_jpen)
        assert abs(float(gn.log_beta) - g["traj"][k, 0]) <= 2e-5 and abs(float(gn.log_gamma) - g["traj"][k, 1]) <= 2e-5, k
        assert np.allclose([float(reg), float(beta), float(gamma)], g["regs"][k], rtol=1e-4)
    assert int(gn.step_count) == int(g["step_count"])
    assert np.allclose([float(gn.L0_anti), float(gn.L0_jpen)], g["L0"], rtol=1e-5)
    assert np.allclose([float(gn.ema_g_anti), float(gn.ema_g_jpen)], g["ema"], rtol=1e-4)
    assert set(gn.state_dict()) == {"log_beta", "log_gamma", "step_count", "L0_anti", "L0_jpen", "ema_g_anti", "ema_g_jpen"}


def test_every_package_module_imports_and_bcf_ppo_mirrors_refuse_cpu():
    """All product modules import on a CPU-only box (syntax / dependency check); the BCF and PPO mirrors have the
    reference's names and refuse CPU tensors instead of falling back."""
    import importlib
    import pkgutil

    import modril_b200

    for m in pkgutil.walk_packages(modril_b200.__path__, "modril_b200."):
        importlib.import_module(m.name)
    from modril_b200.bcf import FlowPolicy, VectorNetwork
    from modril_b200.ppo import compute_advantages

    net = VectorNetwork(11, 3, 64, 32)
    assert list(net.state_dict()) == ["time_embed.0.weight", "time_embed.0.bias", "time_embed.2.weight", "time_embed.2.bias",
                                      "state_embed.0.weight", "state_embed.0.bias", "action_embed.0.weight",
                                      "action_embed.0.bias", "velocity_head.0.weight", "velocity_head.0.bias",
                                      "velocity_head.2.weight", "velocity_head.2.bias"]
    with pytest.raises(RuntimeError):
        FlowPolicy(net).get_action(torch.zeros(2, 11))
    with pytest.raises(RuntimeError):
        compute_advantages(torch.zeros(3, 2, 1), torch.zeros(3, 2, 1))


def test_ppo_policy_surface_without_a_gpu():
    """DistActorCritic mirror: the reference policy's state_dict keys / shapes / init, one flat buffer, no CPU fallback;
    PPO's flags are the reference's."""
    import argparse

    from modril_b200.ppo import PPO, DistActorCritic
    from oracle import ppo_closed_form as pcf

    m = DistActorCritic(17, 6, 64, 3)
    spec = pcf.ACSpec(17, 6, 64, 3)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(n, tuple(s)) for n, s in spec.shapes()]
    assert m.flat_params().numel() == spec.n_params()
    off = 0
    for p in m.parameters():
        assert p.data_ptr() == m.flat_params().data_ptr() + 4 * off
        off += p.numel()
    new = pcf.ac_make_params(spec, 2)
    m.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in pcf.ac_unpack(spec, new).items()})
    assert np.array_equal(m.flat_params().numpy(), new)
    with pytest.raises(RuntimeError):
        m.get_value(torch.zeros(2, 17))
    with pytest.raises(RuntimeError):
        m.ppo_loss(torch.zeros(2, 17), torch.zeros(2, 6), *(torch.zeros(2) for _ in range(4)))
    parser = argparse.ArgumentParser()
    PPO().get_add_args(parser)
    a = parser.parse_args([])
    assert (a.num_epochs, a.num_mini_batch, a.clip_param, a.entropy_coef, a.value_loss_coef, a.max_grad_norm, a.lr, a.eps,
            a.linear_lr_decay) == (4, 4, 0.2, 0.01, 0.5, 0.5, 1e-3, 1e-12, True)


def test_ppo_oracle_tie_and_clip_rules_match_autograd():
    """The oracle's hand-written backward follows autograd on the non-smooth points: ratio exactly inside / outside the
    clip range, value clip active, zero advantages (torch.min / torch.max ties)."""
    from oracle import ppo_closed_form as pcf

    spec = pcf.ACSpec(5, 2, 16, 1)
    flat = pcf.ac_make_params(spec, 4).astype(np.float64)
    g = np.random.default_rng(0)
    B = 64
    obs, act = g.standard_normal((B, 5)), g.standard_normal((B, 2)) * 0.3
    _, logp, _ = pcf.ac_evaluate(spec, flat, obs, act)
    old_logp = logp + g.choice([-0.5, 0.0, 0.5], size=(B, 1))        # ratio = exp(+-0.5) or exactly 1
    adv = g.choice([-1.0, 0.0, 1.0], size=(B, 1))
    value = pcf.ac_forward(spec, flat, obs)[0]
    old_value = value + g.choice([-0.5, 0.0, 0.5], size=(B, 1))
    ret = value + g.standard_normal((B, 1))
    losses, grad = pcf.ppo_loss_and_grad(spec, flat, obs, act, old_logp, adv, ret, old_value)
    t = torch.tensor(flat, requires_grad=True)
    P, off = {}, 0
    for n, s in spec.shapes():
        k = int(np.prod(s))
        P[n] = t[off:off + k].view(*s)
        off += k
    o = torch.tensor(obs)
    v = torch.tanh(o @ P["critic.net.0.weight"].T + P["critic.net.0.bias"]) @ P["critic_head.weight"].T + P["critic_head.bias"]
    mean = torch.tanh(o @ P["actor.net.0.weight"].T + P["actor.net.0.bias"]) @ P["dist.fc_mean.weight"].T + P["dist.fc_mean.bias"]
    d = torch.distributions.Normal(mean, P["dist.logstd"].expand_as(mean).exp())
    lp, ent = d.log_prob(torch.tensor(act)).sum(-1, keepdim=True), d.entropy().sum(-1)
    ratio, A = torch.exp(lp - torch.tensor(old_logp)), torch.tensor(adv)
    al = -torch.min(ratio * A, torch.clamp(ratio, 0.8, 1.2) * A).mean(0)
    vpc = torch.tensor(old_value) + (v - torch.tensor(old_value)).clamp(-0.2, 0.2)
    vl = 0.5 * torch.max((v - torch.tensor(ret)).pow(2), (vpc - torch.tensor(ret)).pow(2)).mean()
    (vl * 0.5 + al - ent.mean() * 0.01).sum().backward()
    assert abs(losses[0] - float(vl * 0.5 + al - ent.mean() * 0.01)) < 1e-12
    assert np.abs(grad - t.grad.numpy()).max() < 1e-12


def test_reference_sweep_yaml_keys_parse_with_our_flags():
    """Every parameter of the reference's flowril sweep configs (configs/sine/flowril*.yaml, fixture written by
    oracle/gen_golden.py::gen_yaml_keys) that the reward module or the PPO updater owns is accepted by our
    get_add_args with the YAML's value; the rest are rl-toolkit's general run flags (rlf/args.py, base_il.py), which stay
    with the framework."""
    import argparse
    import json
    import os

    from modril_b200.flowril.flowril import FlowMatchingEstimation
    from modril_b200.ppo import PPO

    framework = {"seed", "alg", "prefix", "env-name", "cuda", "eval-interval", "traj-load-path", "eval-num-processes",
                 "num-eval", "num-env-steps"}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flowril_yaml_keys.json")
    configs = json.load(open(path))
    assert len(configs) >= 5
    for name, params in configs.items():
        parser = argparse.ArgumentParser()
        FlowMatchingEstimation().get_add_args(parser)
        PPO().get_add_args(parser)
        argv = []
        for k, v in params.items():
            if k in framework:
                continue
            val = v["value"] if "value" in v else v["values"][0]
            argv += [f"--{k}", str(val)]
        ns = parser.parse_args(argv)            # unknown flag -> SystemExit
        assert ns.reward_type in ("raw", "norm", "exp", "clip", "smooth"), name
        if "2fs" in name:
            assert ns.option == "2fs" and ns.flow_path.endswith("sine_fm.pt")


def test_free_hidden_dim_pad_plan_places_every_true_entry_in_a_zero_padded_image():
    """Trunk widths other than the kernels' 128 / 256 run zero-padded (networks.py::_pad_plan): every entry of the true
    flat vector (state_dict order, reference shapes) lands at its [row, column] inside the padded layer blocks, the
    rest of the image stays zero, and gathering a padded gradient returns the true layout."""
    import torch

    from modril_b200.flowril.networks import ConditionalVNet, SharedVNet

    for cls, kw in ((SharedVNet, dict(s_dim=5, a_dim=2, hidden=100, depth=3)),
                    (ConditionalVNet, dict(s_dim=4, a_dim=3, hidden=192, depth=4)),
                    (SharedVNet, dict(s_dim=3, a_dim=1, hidden=7, depth=2))):
        net = cls(**kw)
        idx, total, image, _ = net._pad_plan()
        H, Hp, d_in = net.hidden, net.native_hidden(), net.a_dim + net.s_dim + 1
        assert Hp in (128, 256) and Hp >= H
        assert total == Hp * d_in + Hp + (net.depth - 1) * (Hp * Hp + Hp) + net.n_heads * (net.a_dim * Hp + net.a_dim)
        assert idx.unique().numel() == idx.numel() == net.flat_params().numel() and int(idx.max()) < total
        image.index_copy_(0, idx, net.flat_params().detach())
        if net.depth >= 2:
            off = Hp * d_in + Hp
            block = image[off:off + Hp * Hp].view(Hp, Hp)
            name = [k for k, _ in net.named_parameters() if k.endswith("2.weight")][0]
            assert torch.equal(block[:H, :H], dict(net.named_parameters())[name].detach())
            assert float(block[H:].abs().sum()) == 0.0 and float(block[:, H:].abs().sum()) == 0.0
        assert float(image.abs().sum()) == pytest.approx(float(net.flat_params().detach().abs().sum()), rel=1e-6)
        assert torch.equal(net.from_native_grad(image), net.flat_params().detach())
    native = SharedVNet(5, 2, hidden=256, depth=4)
    assert native._pad_plan() is None and native.native_hidden() == 256
    with pytest.raises(ValueError):
        SharedVNet(5, 2, hidden=300, depth=2).native_hidden()
