"""GPU parity of the PPO policy / update kernels (frl_acnet_forward / _act / _evaluate / frl_ppo_loss_grad through the C
ABI) against outputs of the reference's own MLPBasic / DiagGaussian classes (tests/golden/ppo_*.npz, ppotrain_*.npz) and
the CPU oracle (oracle/ppo_closed_form.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import ppo_closed_form as pcf
from oracle import synth
from oracle.gen_golden import ppo_inputs
from tests.gpu_common import dev

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def build_policy(spec, flat, **kw):
    from modril_b200.ppo import DistActorCritic

    m = DistActorCritic(spec.obs_dim, spec.a_dim, spec.hidden, spec.layers, **kw).to("cuda:0")
    m.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in pcf.ac_unpack(spec, flat).items()})
    return m


def _rel_l2(got, ref):
    ref = np.asarray(ref, np.float64)
    return np.sqrt(((np.asarray(got, np.float64) - ref) ** 2).sum()) / np.sqrt((ref ** 2).sum())


def _close(got, ref, rel):
    ref = np.asarray(ref, np.float64)
    return np.abs(np.asarray(got, np.float64) - ref).max() <= rel * max(1.0, np.abs(ref).max())


def test_state_dict_keys_and_init_match_reference_policy():
    from modril_b200.ppo import DistActorCritic

    spec = pcf.ACSpec(11, 3)
    m = DistActorCritic(11, 3)
    sd = m.state_dict()
    assert [(k, tuple(v.shape)) for k, v in sd.items()] == [(n, tuple(s)) for n, s in spec.shapes()]
    assert all(float(b.abs().max()) == 0.0 for n, b in sd.items() if n.endswith("bias") or n == "dist.logstd")
    w = sd["actor.net.2.weight"].double()
    assert torch.allclose(w @ w.T, 2.0 * torch.eye(64, dtype=torch.float64), atol=1e-5)      # orthogonal, gain sqrt(2)
    with pytest.raises(RuntimeError):
        m.get_value(torch.zeros(2, 11))                                                       # CPU tensors: no fallback


@pytest.fixture(params=["fused", "gemm"])
def kernel_path(request, monkeypatch):
    """frl_ppo_loss_grad has two CUDA paths: the fused shared-memory kernel (nets that fit; the default) and the per-layer
    GEMM path (FRL_PPO_GEMM=1 forces it).  Both must meet the same bar."""
    if request.param == "gemm":
        monkeypatch.setenv("FRL_PPO_GEMM", "1")
    else:
        monkeypatch.delenv("FRL_PPO_GEMM", raising=False)
    return request.param


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ppo_*.npz"))), ids=os.path.basename)
def test_policy_loss_grad_vs_reference(path, kernel_path):
    g = np.load(path)
    s_dim, a_dim, hidden, layers, B, seed = (int(x) for x in g["meta"])
    clip, vcoef, ecoef = (float(x) for x in g["coefs"])
    spec = pcf.ACSpec(s_dim, a_dim, hidden, layers)
    flat = pcf.ac_make_params(spec, seed)
    obs, action, old_logp, adv, ret, old_value, noise = ppo_inputs(spec, flat, seed, B)
    m = build_policy(spec, flat)
    mean, value = m(dev(obs))
    assert _close(mean.cpu().numpy(), g["mean_f64"], 2e-5) and _close(value.cpu().numpy(), g["value_f64"], 2e-5)
    assert torch.equal(m.get_value(dev(obs)), value)
    ev = m.evaluate_actions(dev(obs), None, None, None, dev(action))
    assert ev["value"].shape == (B, 1) and ev["log_prob"].shape == (B, 1) and ev["ent"].shape == (B,)
    assert _close(ev["log_prob"].cpu().numpy(), g["logp_f64"], 2e-5)
    assert _close(ev["ent"].cpu().numpy(), g["ent_f64"], 1e-5) and torch.equal(ev["value"], value)
    # get_action with the injected draw: same action / log-prob as the oracle; deterministic policy returns the mean
    act = m.get_action(dev(obs), noise=dev(noise))
    _, a_ref, lp_ref = pcf.ac_act(spec, flat, obs, noise)
    assert _close(act["action"].cpu().numpy(), a_ref, 2e-5) and _close(act["action_log_probs"].cpu().numpy(), lp_ref, 5e-5)
    det = build_policy(spec, flat, deterministic_policy=True).get_action(dev(obs))
    assert torch.equal(det["action"], mean)

    args = (dev(obs), dev(action), dev(old_logp), dev(adv), dev(ret), dev(old_value))
    kw = dict(clip_param=clip, value_loss_coef=vcoef, entropy_coef=ecoef)
    losses = m.ppo_loss(*args, **kw).cpu().numpy()
    assert np.allclose(losses, g["loss_f64"], rtol=2e-5, atol=2e-6), (losses, g["loss_f64"])
    grad = m.flat_grad().cpu().numpy()
    ref = g["grad_f64"].astype(np.float64)
    assert _rel_l2(grad, ref) <= 1e-4
    off = 0
    for name, shape in spec.shapes():
        n = int(np.prod(shape))
        r = ref[off:off + n]
        if np.abs(r).max() > 0:
            assert _rel_l2(grad[off:off + n], r) <= 3e-4, name
        off += n
    # loss only leaves the gradient alone; accumulate adds; grad_scale scales
    before = m.flat_grad().clone()
    l2 = m.ppo_loss(*args, grad=False, **kw)
    assert torch.equal(before, m.flat_grad()) and torch.equal(l2.cpu(), torch.from_numpy(losses))
    m.ppo_loss(*args, accumulate=True, grad_scale=0.5, **kw)
    assert _rel_l2(m.flat_grad().cpu().numpy(), 1.5 * ref) <= 1e-4
    # the fused minibatch gather: a shuffled index over a larger "rollout" gives the same result as pre-gathered rows
    perm = torch.randperm(B, device="cuda", generator=torch.Generator("cuda").manual_seed(seed))
    inv = torch.argsort(perm)
    shuffled = tuple(x[perm].contiguous() for x in args)
    l3 = m.ppo_loss(*shuffled, indices=inv, **kw)
    assert torch.equal(l3.cpu(), torch.from_numpy(losses)) and _rel_l2(m.flat_grad().cpu().numpy(), ref) <= 1e-4


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ppotrain_*.npz"))), ids=os.path.basename)
def test_ppo_update_follows_reference_curve(path):
    from modril_b200.ppo import FusedAdam, ppo_update

    g = np.load(path)
    s_dim, a_dim, hidden, layers, rows, nmb, epochs, seed = (int(x) for x in g["meta"])
    spec = pcf.ACSpec(s_dim, a_dim, hidden, layers)
    flat = pcf.ac_make_params(spec, seed)
    obs, action, old_logp, adv, ret, old_value, _ = ppo_inputs(spec, flat, seed, rows)
    m = build_policy(spec, flat)
    opt = FusedAdam([m], lr=float(g["lr"]), eps=1e-12)
    perms = [torch.from_numpy(synth.rng(seed + 100 + e).permutation(rows)).cuda() for e in range(epochs)]
    data = (dev(obs), dev(action), dev(old_logp), dev(old_value), dev(ret), dev(adv))
    # step by step (the curve), on a copy driven by hand
    curve, norms = [], []
    mb = rows // nmb
    for e in range(epochs):
        for k in range(nmb):
            i = perms[e][k * mb:(k + 1) * mb].contiguous()
            curve.append(m.ppo_loss(data[0], data[1], data[2], data[5], data[4], data[3], indices=i))
            opt.step_flat([m.flat_grad()], max_norm=0.5)
            norms.append(opt._norm.clone())
    curve, norms = torch.stack(curve).cpu().numpy(), torch.cat(norms).cpu().numpy()
    assert np.allclose(curve[:12], g["curve"][:12], rtol=2e-3, atol=2e-4), np.abs(curve[:12] - g["curve"][:12]).max()
    assert np.allclose(curve, g["curve"], rtol=5e-2, atol=5e-3), np.abs(curve - g["curve"]).max()
    assert np.allclose(norms[:12], g["gradnorm"][:12], rtol=5e-3)
    assert _rel_l2(m.flat_params().cpu().numpy(), g["final_params"]) <= 2e-2
    # the same thing through ppo_update (the PPO.update mirror): identical parameters and the reference's log_vals
    m2 = build_policy(spec, flat)
    opt2 = FusedAdam([m2], lr=float(g["lr"]), eps=1e-12)
    logs = ppo_update(m2, opt2, *data, num_epochs=epochs, num_mini_batch=nmb, perms=perms)
    assert torch.equal(m2.flat_params(), m.flat_params())
    n = epochs * nmb
    assert abs(logs["value_loss"] - g["curve"][:, 1].mean()) <= 5e-2 * abs(g["curve"][:, 1].mean())
    assert abs(logs["dist_entropy"] - g["curve"][:, 3].mean()) <= 1e-2 * abs(g["curve"][:, 3].mean())
    assert logs["policy_update_data"] == float(nmb * n)
    # without perms the sampler draws its own permutation and still runs
    ppo_update(m2, opt2, *data, num_epochs=1, num_mini_batch=nmb)
    assert torch.isfinite(m2.flat_params()).all()


def test_graphed_update_equals_eager_update():
    """PPOUpdateGraph (one CUDA-graph replay per PPO.update) == ppo_update with the same capturable optimiser, bit for bit,
    over two consecutive updates with different data."""
    from modril_b200.ppo import FusedAdam, PPOUpdateGraph, ppo_update

    spec = pcf.ACSpec(11, 3, 64, 2)
    flat = pcf.ac_make_params(spec, 5)
    rows, nmb, epochs = 1000, 4, 3                      # 1000 // 4 = 250-row minibatches
    pols = [build_policy(spec, flat) for _ in range(2)]
    opts = [FusedAdam([p], lr=1e-3, eps=1e-12, capturable=True) for p in pols]
    graphed = PPOUpdateGraph(pols[1], opts[1], rows, num_epochs=epochs, num_mini_batch=nmb)
    assert torch.equal(pols[1].flat_params().cpu(), torch.from_numpy(flat)) and opts[1].step_count() == 0   # warm-up undone
    for rnd in range(2):
        obs, action, old_logp, adv, ret, old_value, _ = ppo_inputs(spec, flat, 50 + rnd, rows)
        data = (dev(obs), dev(action), dev(old_logp), dev(old_value), dev(ret), dev(adv))
        perms = [torch.from_numpy(synth.rng(60 + 10 * rnd + e).permutation(rows)).cuda() for e in range(epochs)]
        la = ppo_update(pols[0], opts[0], *data, num_epochs=epochs, num_mini_batch=nmb, perms=perms)
        lb = graphed(*data, perms=perms)
        assert la == lb
        assert torch.equal(pols[0].flat_params(), pols[1].flat_params())
        assert opts[0].step_count() == opts[1].step_count() == (rnd + 1) * epochs * nmb
        # the policy is usable right after a replay (weights re-packed): same values from both
        assert torch.equal(pols[0].get_value(data[0]), pols[1].get_value(data[0]))
        if rnd == 0:
            # a larger eager call outgrows the workspace the graph was captured with: the old buffer must stay alive
            big = ppo_inputs(spec, flat, 99, 40000)[:6]
            pols[1].ppo_loss(*(dev(x) for x in big), grad=False)
    graphed(*data)                                       # own permutations
    assert torch.isfinite(pols[1].flat_params()).all()


def test_ppo_updater_runs_the_whole_consumer_chain_on_the_device():
    """PPO.update (next value -> returns / GAE -> advantages -> minibatches) against the oracle's composition of the same
    steps on the first minibatch, eager and graphed; flags and checkpoint keys of the reference class."""
    import argparse
    from types import SimpleNamespace

    from modril_b200.ppo import PPO

    T, N, S, A = 64, 8, 11, 3
    spec = pcf.ACSpec(S, A, 64, 2)
    flat = pcf.ac_make_params(spec, 31)
    g = synth.rng(32)
    obs = g.standard_normal((T + 1, N, S)).astype(np.float32)
    noise = g.standard_normal((T * N, A)).astype(np.float32)
    value, action, logp = (x.astype(np.float32) for x in pcf.ac_act(spec, flat, obs[:-1].reshape(T * N, S), noise))
    rewards = g.standard_normal((T, N, 1)).astype(np.float32)
    masks = (g.random((T + 1, N, 1)) > 0.05).astype(np.float32)
    parser = argparse.ArgumentParser()
    updater = PPO()
    updater.get_add_args(parser)
    args = parser.parse_args(["--linear-lr-decay", "false"])
    assert (args.num_epochs, args.num_mini_batch, args.clip_param, args.max_grad_norm, args.eps) == (4, 4, 0.2, 0.5, 1e-12)
    args.gamma, args.gae_lambda, args.use_gae, args.num_steps, args.num_processes = 0.99, 0.95, True, T, N

    def storage():
        vp = np.concatenate([value.reshape(T, N, 1), np.zeros((1, N, 1), np.float32)])
        return SimpleNamespace(obs=dev(obs), actions=dev(action.reshape(T, N, A)), action_log_probs=dev(logp.reshape(T, N, 1)),
                               rewards=dev(rewards), value_preds=dev(vp), returns=torch.zeros(T + 1, N, 1, device="cuda"),
                               masks=dev(masks), bad_masks=dev(np.ones_like(masks)))

    perms = [torch.from_numpy(synth.rng(40 + e).permutation(T * N)).cuda() for e in range(4)]
    finals = []
    for graph in (False, True):
        pol = build_policy(spec, flat)
        updater = PPO()
        updater.init(pol, args)
        st = storage()
        logs = updater.update(st, perms=perms, graph=graph)
        finals.append(pol.flat_params().clone())
        assert set(logs) == {"value_loss", "actor_loss", "dist_entropy", "policy_update_data"}
        assert updater.opt.step_count() == 16
    assert torch.equal(finals[0], finals[1])
    # oracle composition: next value, GAE, normalised advantages, then the first minibatch of epoch 0
    nv = pcf.ac_forward(spec, flat, obs[-1])[0].astype(np.float32)
    vp = np.concatenate([value.reshape(T, N, 1), np.zeros((1, N, 1), np.float32)])
    returns, vp = pcf.compute_returns(rewards, vp, masks, np.ones_like(masks), nv, 0.99, 0.95, True, False)
    assert np.abs(st.returns[:-1].cpu().numpy() - returns[:-1]).max() <= 1e-5
    adv = pcf.compute_advantages(returns, vp)
    i = perms[0][:T * N // 4].cpu().numpy()
    flat_rows = lambda x: x.reshape(T * N, -1)[i]
    losses, _ = pcf.ppo_loss_and_grad(spec, flat, flat_rows(obs[:-1]), flat_rows(action), flat_rows(logp), flat_rows(adv),
                                      flat_rows(returns[:-1]), flat_rows(vp[:-1]))
    pol = build_policy(spec, flat)
    got = pol.ppo_loss(st.obs[:-1].reshape(T * N, S), st.actions.reshape(T * N, A), st.action_log_probs.reshape(-1),
                       dev(adv).reshape(-1), st.returns[:-1].reshape(-1), st.value_preds[:-1].reshape(-1),
                       indices=perms[0][:T * N // 4].contiguous()).cpu().numpy()
    assert np.allclose(got, losses, rtol=5e-5, atol=5e-6), (got, losses)
    # lr decay (base_net_algo.py:78-82) and the checkpoint key
    args.linear_lr_decay, args.num_env_steps = True, 10 * T * N
    updater.init(pol, args)
    updater.pre_update(5)
    assert abs(updater.opt.param_groups[0]["lr"] - 0.5e-3) < 1e-12
    saved = {}
    updater.save(SimpleNamespace(save_key=saved.__setitem__))
    assert "actor_opt" in saved and "state" in saved["actor_opt"]


@pytest.mark.parametrize("hidden", [64, 128, 256])
def test_large_ragged_batch_matches_oracle_and_is_deterministic(hidden, kernel_path):
    spec = pcf.ACSpec(17, 6, hidden, 2)          # 64 / 128: fused kernel with several tiles per CTA; 256: GEMM path
    flat = pcf.ac_make_params(spec, 77)
    B = 4099 if hidden == 256 else 20011
    obs, action, old_logp, adv, ret, old_value, _ = ppo_inputs(spec, flat, 770, B)
    m = build_policy(spec, flat)
    args = [dev(x) for x in (obs, action, old_logp, adv, ret, old_value)]
    l1 = m.ppo_loss(*args).cpu().numpy()
    g1 = m.flat_grad().clone()
    l2 = m.ppo_loss(*args).cpu().numpy()
    assert np.array_equal(l1, l2) and torch.equal(g1, m.flat_grad())
    losses, grad = pcf.ppo_loss_and_grad(spec, flat, obs, action, old_logp, adv, ret, old_value)
    assert np.allclose(l1, losses, rtol=2e-5, atol=2e-6)
    assert _rel_l2(g1.cpu().numpy(), grad) <= 1e-4
    # data-parallel halves: mean_divisor = global rows, gradients and loss shares add up to the full batch
    h = B // 2
    la = m.ppo_loss(*(x[:h].contiguous() for x in args), mean_divisor=B).cpu().numpy()
    lb = m.ppo_loss(*(x[h:].contiguous() for x in args), mean_divisor=B, accumulate=True).cpu().numpy()
    assert _rel_l2(m.flat_grad().cpu().numpy(), grad) <= 1e-4
    assert np.allclose((la + lb)[:3], losses[:3], rtol=2e-5, atol=2e-6)


def test_bad_arguments_raise():
    from modril_b200.ppo import DistActorCritic

    with pytest.raises(ValueError):
        DistActorCritic(4, 2, ppo_hidden_dim=100).to("cuda:0").native()         # hidden not a multiple of 16
    m = DistActorCritic(4, 2, 32, 1).to("cuda:0")
    z = torch.zeros(3, device="cuda")
    with pytest.raises(ValueError):
        m.get_value(torch.zeros(3, 5, device="cuda"))
    with pytest.raises(ValueError):
        m.ppo_loss(torch.zeros(3, 4, device="cuda"), torch.zeros(3, 2, device="cuda"), z, z, z, torch.zeros(2, device="cuda"))
    with pytest.raises(TypeError):
        m.ppo_loss(torch.zeros(3, 4, device="cuda"), torch.zeros(3, 2, device="cuda"), z, z, z, z,
                   indices=torch.zeros(3, dtype=torch.int32, device="cuda"))
    assert m.get_value(torch.zeros(0, 4, device="cuda")).shape == (0, 1)
