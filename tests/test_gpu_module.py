"""The reward-module mirror (FlowMatchingEstimation hooks) on the GPU against the oracle, with a fake rollout storage.
Reference call sites: flowril.py:175-230 (_get_reward), :232-350 (_update_reward_func), base_irl.py:37-72 (update)."""
import argparse
import os
import copy

import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import synth
from tests.gpu_common import assert_close

pytestmark = pytest.mark.gpu


class Space:
    def __init__(self, n):
        self.shape = (n,)


class Policy:
    def __init__(self, s_dim, a_dim):
        self.obs_space, self.action_space = Space(s_dim), Space(a_dim)


class FakeStorage:
    """The slice of rlf RolloutStorage the module touches (rollout_storage.py:243-323)."""

    def __init__(self, T, N, s_dim, a_dim, seed, device):
        g = synth.rng(seed)
        self.obs = torch.from_numpy(g.standard_normal((T + 1, N, s_dim)).astype(np.float32)).to(device)
        self.actions = torch.from_numpy(np.clip(g.standard_normal((T, N, a_dim)), -1, 1).astype(np.float32)).to(device)
        m = (g.random((T + 1, N, 1)) > 0.1).astype(np.float32)
        self.masks = torch.from_numpy(m).to(device)
        self.rewards = torch.zeros(T, N, 1, device=device)
        self.T, self.N = T, N

    def get_obs(self, step):
        return self.obs[step]

    def get_generator(self, advantages=None, num_mini_batch=None, mini_batch_size=None, **kw):
        B = self.T * self.N
        perm = torch.randperm(B, generator=torch.Generator().manual_seed(1))
        flat_s = self.obs[:-1].reshape(B, -1)
        flat_a = self.actions.reshape(B, -1)
        for i in range(0, B - mini_batch_size + 1, mini_batch_size):
            idx = perm[i:i + mini_batch_size].to(flat_s.device)
            yield {"state": flat_s[idx], "other_state": {}, "action": flat_a[idx]}


def make_args(**kw):
    d = dict(option="scrf", hidden_dim=128, flow_depth=4, flow_path=None, disc_lr=1e-3, disc_grad_pen=1.0,
             expert_loss_rate=1.0, agent_loss_rate=1.0, enable_loss_anti=False, enable_loss_stable=False,
             params_autotune=False, loss_anti=1e-4, loss_stable=1e-6, finetune_ve=False, reward_type="raw",
             flowril_state_norm=False, flowril_reward_norm=True, n_flowril_epochs=1, gamma=0.99, num_steps=6,
             num_processes=8, traj_batch_size=16, device=torch.device("cuda:0"), flowril_precision="fp32",
             flowril_batched_scoring=True, env_name="Hopper-v3")
    d.update(kw)
    return argparse.Namespace(**d)


def make_module(args, s_dim=11, a_dim=3, seed=0):
    from modril_b200.flowril.flowril import FlowMatchingEstimation
    torch.manual_seed(seed)
    g = synth.rng(99)
    obs = torch.from_numpy(g.standard_normal((64, s_dim)).astype(np.float32))
    args.expert_dataset = {"obs": obs, "actions": torch.from_numpy(synth.expert_actions(5, obs.numpy(), a_dim))}
    mod = FlowMatchingEstimation()
    mod.init(Policy(s_dim, a_dim), args)
    return mod


def test_get_reward_matches_oracle_with_return_normalisation():
    from modril_b200.flowril import _rademacher
    args = make_args()
    mod = make_module(args)
    T, N = args.num_steps, args.num_processes
    storage = FakeStorage(T, N, 11, 3, 7, args.device)
    spec = cf.NetSpec(11, 3, 128, 4, 2)
    flat = mod.flow_net.net.flat_params().cpu().numpy().astype(np.float64)
    probe = _rademacher((N, 3), device=args.device).cpu().numpy()
    raw = np.zeros((T, N, 1), np.float32)
    for t in range(T):
        s, a = storage.obs[t].cpu().numpy(), storage.actions[t].cpu().numpy()
        le = cf.log_prob(spec, flat, s, a, "expert", probe, 32)
        la = cf.log_prob(spec, flat, s, a, "agent", probe, 32)
        raw[t, :, 0] = (le - la).astype(np.float32)
    ref, _, _ = cf.normalise_rewards(raw, storage.masks.cpu().numpy(), args.gamma)
    got = np.stack([mod._get_reward(t, storage, {})[0].cpu().numpy() for t in range(T)])
    assert_close(got, ref, 2e-4, "normalised rewards", 0.03)
    # per-step (reference call shape) and batched scoring are the same numbers, bit for bit
    args2 = make_args(flowril_batched_scoring=False)
    mod2 = make_module(args2)
    got2 = np.stack([mod2._get_reward(t, storage, {})[0].cpu().numpy() for t in range(T)])
    assert np.array_equal(got, got2)
    # disc value (flowril.py:352-365)
    dv = mod._compute_disc_val(storage.obs[0], storage.actions[0]).cpu().numpy()
    assert_close(dv, raw[0, :, 0], 1e-4, "disc val", 0.13)


@pytest.mark.parametrize("reward_type", ["norm", "exp", "clip", "smooth"])
def test_reward_types_through_the_module(reward_type):
    args = make_args(reward_type=reward_type, flowril_reward_norm=False)
    mod = make_module(args)
    storage = FakeStorage(args.num_steps, args.num_processes, 11, 3, 8, args.device)
    raw_args = make_args(reward_type="raw", flowril_reward_norm=False)
    raw_mod = make_module(raw_args)
    for t in range(2):
        got = mod._get_reward(t, storage, {})[0].cpu().numpy()
        raw = raw_mod._get_reward(t, storage, {})[0].cpu().numpy()
        assert_close(got, cf.reward_transform(raw, reward_type), 2e-6, reward_type)


def test_bad_reward_type():
    args = make_args(reward_type="bogus")
    mod = make_module(args)
    storage = FakeStorage(args.num_steps, args.num_processes, 11, 3, 8, args.device)
    with pytest.raises(ValueError):
        mod._get_reward(0, storage, {})


def test_update_matches_oracle_including_role_swap():
    """One _update_reward_func pass: replay the module's random draws (same torch seeds, same order) through the
    float64 oracle.  Role 'expert' is trained on the ROLLOUT batch and role 'agent' on the DEMO batch."""
    args = make_args(num_steps=4, num_processes=8, traj_batch_size=16)
    mod = make_module(args)
    storage = FakeStorage(args.num_steps, args.num_processes, 11, 3, 9, args.device)
    spec = cf.NetSpec(11, 3, 128, 4, 2)
    flat = mod.flow_net.net.flat_params().cpu().numpy().astype(np.float64)
    m, v = np.zeros_like(flat), np.zeros_like(flat)
    # replay: the DataLoader shuffle and the t/noise draws come from torch's global generators
    state_cpu, state_cuda = torch.get_rng_state(), torch.cuda.get_rng_state()
    log_vals = mod._update_reward_func(storage)
    torch.set_rng_state(state_cpu)
    torch.cuda.set_rng_state(state_cuda)
    losses_ref = []
    step = 0
    for demo, roll in zip(mod.expert_train_loader, storage.get_generator(None, mini_batch_size=16)):
        batches = []
        for role, s, a in (("expert", roll["state"], roll["action"]), ("agent", demo["state"], demo["actions"])):
            B = a.shape[0]
            t = torch.rand(B, 1, device=args.device) * (1 - 2e-3) + 1e-3
            noise = torch.distributions.Normal(0., 1.).sample((B, 3))
            batches.append((role, s.cpu().numpy(), a.cpu().numpy(), t.cpu().numpy(), noise.numpy()))
        step += 1
        flat, m, v, losses, _ = cf.train_step(spec, flat, m, v, step, batches, lr=args.disc_lr, max_norm=1.0)
        losses_ref.append(losses)
    assert step == 2
    ref = np.mean(np.array(losses_ref), axis=0)
    assert abs(log_vals["flow_loss"] - ref[0]) <= 1e-4 * max(1, abs(ref[0]))
    assert abs(log_vals["agent_loss"] - ref[1]) <= 1e-4 * max(1, abs(ref[1]))
    got = mod.flow_net.net.flat_params().cpu().numpy()
    assert np.abs(got - flat).max() <= 5e-6


def test_full_update_loop_and_checkpoint_roundtrip():
    args = make_args()
    mod = make_module(args)
    storage = FakeStorage(args.num_steps, args.num_processes, 11, 3, 11, args.device)
    logs = mod.update(storage)
    assert set(logs) >= {"flow_loss", "agent_loss"}
    assert torch.isfinite(storage.rewards).all() and storage.rewards.abs().sum() > 0

    class Ckpt:
        def __init__(self):
            self.d = {}

        def save_key(self, k, v):
            self.d[k] = copy.deepcopy(v)

        def get_key(self, k):
            return self.d[k]

    ck = Ckpt()
    mod.save(ck)
    assert set(ck.d) == {"flow_scrf_opt", "flow_scrf"}
    assert list(ck.d["flow_scrf"].keys())[0] == "net.shared.0.weight"
    args2 = make_args()
    mod2 = make_module(args2, seed=123)
    mod2.load_resume(ck)
    assert torch.equal(mod2.flow_net.net.flat_params(), mod.flow_net.net.flat_params())
    # resumed optimiser continues identically
    l1 = mod._update_reward_func(storage)
    torch.manual_seed(5); s1 = torch.get_rng_state(); c1 = torch.cuda.get_rng_state()
    for mm in (mod, mod2):
        torch.set_rng_state(s1); torch.cuda.set_rng_state(c1)
        mm.flow_net.load_state_dict(ck.d["flow_scrf"]); mm.opt.load_state_dict(ck.d["flow_scrf_opt"])
        mm._update_reward_func(storage)
    assert torch.equal(mod2.flow_net.net.flat_params(), mod.flow_net.net.flat_params())


@pytest.mark.parametrize("autotune", [True, False])
def test_update_with_default_regularisers_matches_oracle(autotune):
    """Default flags of the reference (--enable-loss-anti/-stable true, --params_autotune true, flowril.py:563-567):
    the anti-divergence and Jacobian-stability terms join the update with the warm-up weights
    L_E / (L_reg + 1e-8) * 1e-3 (flowril.py:272-284); with --params_autotune false the fixed --loss-anti /
    --loss-stable weights (flowril.py:266-270).  Replayed through the float64 oracle with the same draws."""
    from modril_b200.flowril import _rademacher
    args = make_args(num_steps=4, num_processes=8, traj_batch_size=16, enable_loss_anti=True, enable_loss_stable=True,
                     params_autotune=autotune)
    mod = make_module(args)
    storage = FakeStorage(args.num_steps, args.num_processes, 11, 3, 9, args.device)
    spec = cf.NetSpec(11, 3, 128, 4, 2)
    flat = mod.flow_net.net.flat_params().cpu().numpy().astype(np.float64)
    m, v = np.zeros_like(flat), np.zeros_like(flat)
    state_cpu, state_cuda = torch.get_rng_state(), torch.cuda.get_rng_state()
    log_vals = mod._update_reward_func(storage)
    torch.set_rng_state(state_cpu)
    torch.cuda.set_rng_state(state_cuda)
    probe = _rademacher((32, 3), device=args.device).cpu().numpy()
    losses_ref, step, w = [], 0, None
    for demo, roll in zip(mod.expert_train_loader, storage.get_generator(None, mini_batch_size=16)):
        batches = []
        for role, s, a in (("expert", roll["state"], roll["action"]), ("agent", demo["state"], demo["actions"])):
            B = a.shape[0]
            t = torch.rand(B, 1, device=args.device) * (1 - 2e-3) + 1e-3
            noise = torch.distributions.Normal(0., 1.).sample((B, 3))
            batches.append((role, s.cpu().numpy(), a.cpu().numpy(), t.cpu().numpy(), noise.numpy()))
        t_mix = torch.rand(32, 1, device=args.device).cpu().numpy()
        warm = ("warmup", 1e-3, 1e-8)
        reg = dict(s=np.concatenate([batches[0][1], batches[1][1]]), a=np.concatenate([batches[0][2], batches[1][2]]),
                   t=t_mix, probe=probe, anti=warm if autotune else 1e-4, stable=warm if autotune else 1e-6)
        step += 1
        flat, m, v, losses, _, (la, ls, wa, wst) = cf.train_step(spec, flat, m, v, step, batches, lr=args.disc_lr,
                                                                  max_norm=1.0, reg=reg)
        losses_ref.append(losses)
        w = (wa, wst)
    assert step == 2
    ref = np.mean(np.array(losses_ref), axis=0)
    assert abs(log_vals["flow_loss"] - ref[0]) <= 1e-4 * max(1, abs(ref[0]))
    assert abs(log_vals["agent_loss"] - ref[1]) <= 1e-4 * max(1, abs(ref[1]))
    got = mod.flow_net.net.flat_params().cpu().numpy()
    assert np.abs(got - flat).max() <= 1e-5, np.abs(got - flat).max()
    # args.loss_anti / args.loss_stable hold the weights last used, like the reference (flowril.py:278, :283)
    assert abs(float(mod.args.loss_anti) - w[0]) <= 1e-3 * w[0]
    assert abs(float(mod.args.loss_stable) - w[1]) <= 1e-3 * w[1]
    if autotune:  # reference quirk: step_count never advances from _compute_flow_loss, warm-up is what runs
        assert int(mod.gradnorm.step_count) == 0


def test_gradnorm_branch_runs_when_step_count_is_past_warmup():
    """The GradNorm2D branch of flowril.py:286-290 (not reachable through the reference's own control flow, see the
    module docstring): beta / gamma weight the regulariser gradients and update() consumes the per-tensor norm sums."""
    args = make_args(num_steps=4, num_processes=8, traj_batch_size=16, enable_loss_anti=True, enable_loss_stable=True,
                     params_autotune=True)
    mod = make_module(args)
    mod.gradnorm.set_steps(mod.gradnorm.warmup_steps + 9)   # next forward() makes it a multiple of update_freq
    mod.gradnorm.L0_anti.fill_(1.0)
    mod.gradnorm.L0_jpen.fill_(1.0)
    storage = FakeStorage(args.num_steps, args.num_processes, 11, 3, 9, args.device)
    lb, lg = float(mod.gradnorm.log_beta.detach()), float(mod.gradnorm.log_gamma.detach())
    before = mod.flow_net.net.flat_params().clone()
    mod._update_reward_func(storage)
    assert int(mod.gradnorm.step_count) == mod.gradnorm.warmup_steps + 11
    assert float(mod.gradnorm.log_beta) != lb and float(mod.gradnorm.log_gamma) != lg
    assert float(mod.gradnorm.ema_g_anti) > 0 and float(mod.gradnorm.ema_g_jpen) > 0
    assert not torch.equal(before, mod.flow_net.net.flat_params())
    assert abs(float(mod.args.loss_anti) - np.exp(lb)) <= 1e-6 * np.exp(lb) or float(mod.args.loss_anti) > 0


def test_2fs_option_runs_and_freezes_expert_net(tmp_path):
    from modril_b200.flowril import FlowMatching
    pre = FlowMatching(11, 3, 128, depth=4)
    path = str(tmp_path / "hopper_fm.pt")
    torch.save(pre.state_dict(), path)
    args = make_args(option="2fs", flow_path=path, flowril_reward_norm=False)
    mod = make_module(args)
    storage = FakeStorage(args.num_steps, args.num_processes, 11, 3, 12, args.device)
    before_e = mod.flow_net_e.net.flat_params().clone()
    before_a = mod.flow_net_a.net.flat_params().clone()
    logs = mod.update(storage)
    assert torch.equal(mod.flow_net_e.net.flat_params(), before_e)          # frozen (flowril.py:101-106)
    assert not torch.equal(mod.flow_net_a.net.flat_params(), before_a)
    assert np.isfinite(logs["flow_loss"]) and np.isfinite(logs["agent_loss"])
    assert torch.isfinite(storage.rewards).all()


@pytest.mark.parametrize("T,N", [(128, 32), (5, 1), (16, 3000), (1, 7)])
def test_device_return_normalisation_vs_oracle(T, N):
    """frl_normalise_rewards (one launch per rollout, state on the device) against the oracle's step-by-step numpy
    restatement of flowril.py:221-228 + RunningMeanStd (float32), over two consecutive rollouts (state carry-over)."""
    from modril_b200.flowril.flowril import DeviceRunningMeanStd
    g = synth.rng(T * 1000 + N)
    rms_dev = DeviceRunningMeanStd("cuda:0")
    rms_ref, ret_ref, ret_dev = None, None, None
    for call in range(2):
        rew = (g.standard_normal((T, N, 1)) * 3 + 0.5).astype(np.float32)
        masks = (g.random((T, N, 1)) > 0.1).astype(np.float32)
        ref, ret_ref, rms_ref = cf.normalise_rewards(rew, masks, 0.99, rms_ref, ret_ref)
        got, ret_dev = rms_dev.normalise(torch.from_numpy(rew).cuda(), torch.from_numpy(masks).cuda(), 0.99, ret_dev)
        assert np.abs(got.cpu().numpy() - ref).max() <= 2e-6 * np.abs(ref).max()
        assert np.abs(ret_dev.cpu().numpy() - ret_ref).max() <= 1e-6 * max(1.0, np.abs(ret_ref).max())
        assert abs(rms_dev.var[0] - rms_ref.var[0]) <= 2e-6 * rms_ref.var[0]
        assert abs(rms_dev.mean[0] - rms_ref.mean[0]) <= 2e-6 * max(1.0, abs(rms_ref.mean[0]))
        assert rms_dev.count == rms_ref.count


@pytest.mark.parametrize("path", sorted(__import__("glob").glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "retnorm_*.npz"))),
                         ids=os.path.basename)
def test_device_return_normalisation_vs_reference_goldens(path):
    """Row a10 against the REFERENCE: frl_normalise_rewards vs the committed outputs of the reference's own _get_reward
    (flowril.py:210-230) + RunningMeanStd (running_mean_std.py:3-35), over consecutive rollouts (oracle/gen_golden.py)."""
    from modril_b200.flowril.flowril import DeviceRunningMeanStd
    from tests.test_oracle_vs_golden import _retnorm_inputs
    g = np.load(path)
    rewards, masks = _retnorm_inputs(g)
    rms_dev, ret_dev = DeviceRunningMeanStd("cuda:0"), None
    for ro in range(rewards.shape[0]):
        got, ret_dev = rms_dev.normalise(torch.from_numpy(rewards[ro]).cuda(), torch.from_numpy(masks[ro]).cuda(),
                                         float(g["gamma"]), ret_dev)
        ref = g["normalised"][ro]
        assert np.abs(got.cpu().numpy() - ref).max() <= 2e-6 * np.abs(ref).max()
        assert abs(rms_dev.var[0] - g["rms_var"][ro]) <= 2e-6 * g["rms_var"][ro]
        assert abs(rms_dev.mean[0] - g["rms_mean"][ro]) <= 2e-6 * max(1.0, abs(g["rms_mean"][ro]))
        assert rms_dev.count == float(g["rms_count"][ro])
    assert np.abs(ret_dev.cpu().numpy() - g["final_returns"]).max() <= 1e-6 * max(1.0, np.abs(g["final_returns"]).max())


def test_standalone_flowril_chain_reward_module_then_device_ppo():
    """FLOWRIL() without rl-toolkit = [FlowMatchingEstimation, modril_b200.ppo.PPO]: one ``update(storage)`` trains the
    reward model, writes the (normalised) rewards for every rollout step and runs the PPO consumer on them."""
    from modril_b200.flowril.flowril import FLOWRIL, HAVE_RLF
    from modril_b200.ppo import DistActorCritic

    if HAVE_RLF:
        pytest.skip("rl-toolkit is importable: FLOWRIL is the real NestedAlgo")
    args = make_args(num_steps=16, num_processes=8, traj_batch_size=16, lr=1e-3, eps=1e-12, gae_lambda=0.95, use_gae=True,
                     num_epochs=2, num_mini_batch=4, clip_param=0.2, value_loss_coef=0.5, entropy_coef=0.01,
                     max_grad_norm=0.5, linear_lr_decay=False)
    T, N, s_dim, a_dim = args.num_steps, args.num_processes, 11, 3
    g = synth.rng(99)
    obs = torch.from_numpy(g.standard_normal((64, s_dim)).astype(np.float32))
    args.expert_dataset = {"obs": obs, "actions": torch.from_numpy(synth.expert_actions(5, obs.numpy(), a_dim))}
    torch.manual_seed(0)
    policy = DistActorCritic(s_dim, a_dim).to(args.device)
    policy.obs_space, policy.action_space = Space(s_dim), Space(a_dim)
    algo = FLOWRIL()
    algo.init(policy, args)
    storage = FakeStorage(T, N, s_dim, a_dim, 7, args.device)
    out = policy.get_action(storage.obs[:-1].reshape(T * N, s_dim))
    storage.actions = out["action"].reshape(T, N, a_dim)
    storage.action_log_probs = out["action_log_probs"].reshape(T, N, 1)
    storage.value_preds = torch.cat([out["value"].reshape(T, N, 1), torch.zeros(1, N, 1, device=args.device)])
    storage.returns = torch.zeros(T + 1, N, 1, device=args.device)
    storage.bad_masks = torch.ones_like(storage.masks)
    before = policy.flat_params().clone()
    reward_before = algo.modules[0].flow_net.net.flat_params().clone()
    for it in range(2):
        algo.pre_update(it)
        logs = algo.update(storage, args)
        assert {"flow_loss", "agent_loss", "value_loss", "actor_loss", "dist_entropy"} <= set(logs), logs
        assert all(np.isfinite(v) for v in logs.values())
        assert float(storage.rewards.abs().sum()) > 0 and torch.isfinite(storage.returns).all()
    assert not torch.equal(before, policy.flat_params()) and torch.isfinite(policy.flat_params()).all()
    assert not torch.equal(reward_before, algo.modules[0].flow_net.net.flat_params())
    assert algo.modules[1].opt.step_count() == 2 * args.num_epochs * args.num_mini_batch
    # freeze_policy (nested_algo.py:99-101): only the reward module updates
    frozen = policy.flat_params().clone()
    args.freeze_policy = True
    algo.update(storage, args)
    assert torch.equal(frozen, policy.flat_params())


# ---- 2fs: batched rollout scoring with the probes drawn in the reference's RNG order (csrc/probe.cu) -----------------
def test_probe_replay_reproduces_the_sequence_of_randint_like_calls():
    """One launch == the T x 2 x n_steps small ``randint_like`` draws of the reference (pretrain.py:90 under flowril.py:186-189),
    values AND generator state afterwards."""
    args = make_args(option="2fs", flowril_reward_norm=False)
    mod = make_module(args)
    dev = args.device
    gen = torch.cuda.default_generators[dev.index]
    for T, N, n_steps in ((3, 8, 5), (2, 1, 32), (1, 700, 2)):
        torch.cuda.manual_seed(1234 + T)
        torch.rand(10, device=dev)                      # a non-zero starting offset
        state = gen.get_state()
        want = mod._draw_loop(T, N, n_steps)
        end = gen.get_offset()
        gen.set_state(state)
        got = mod._draw_replay(T, N, n_steps)
        assert gen.get_offset() == end
        for w, g in zip(want, got):
            assert torch.equal(w, g)
            assert set(torch.unique(g).tolist()) <= {-1.0, 1.0}
    assert type(mod)._replay_ok.get(dev.index, True)


@pytest.mark.parametrize("reward_type", ["raw", "smooth"])
def test_2fs_batched_rollout_scoring_equals_per_step_scoring(reward_type):
    """The whole rollout in two launches (one per net) == T per-step calls, bit for bit, when the global generator starts
    from the same state: same probes (replayed), same kernels, and the same reward transform kernel."""
    rewards = {}
    for batched in (True, False):
        args = make_args(option="2fs", flowril_reward_norm=False, reward_type=reward_type, flowril_batched_scoring=batched,
                         num_steps=5, num_processes=6)
        mod = make_module(args)
        storage = FakeStorage(args.num_steps, args.num_processes, 11, 3, 21, args.device)
        torch.cuda.manual_seed(77)
        if batched:
            type(mod)._replay_ok.pop(args.device.index, None)   # exercise the self-check too (it restores the generator)
        rows = [mod._get_reward(t, storage, {})[0].clone() for t in range(args.num_steps)]
        rewards[batched] = torch.stack(rows)
        assert rewards[batched].shape == (args.num_steps, args.num_processes, 1)
    assert torch.equal(rewards[True], rewards[False])


def test_stand_in_update_loop_logs_like_base_irl():
    """The stand-in of ``BaseIRLAlgo.update`` (used when rl-toolkit is absent) keeps base_irl.py:57-67's logging:
    per-process running reward sums, pushed to ``ep_log_vals`` when an episode ends, reported as ``culm_irl_reward``."""
    from modril_b200.flowril.flowril import HAVE_RLF

    if HAVE_RLF:
        pytest.skip("rl-toolkit is importable: the real loop is tested in test_gpu_plugin_loop.py")
    args = make_args(num_steps=7, num_processes=4, log_smooth_len=100)
    mod = make_module(args)
    storage = FakeStorage(args.num_steps, args.num_processes, 11, 3, 31, args.device)
    storage.masks[2, 1] = 0.0
    storage.masks[5, 1] = 0.0
    storage.masks[6, 3] = 0.0
    logs = mod.update(storage)
    r = storage.rewards.detach().cpu().numpy()[:, :, 0]
    m = storage.masks.detach().cpu().numpy()[:args.num_steps, :, 0]
    culm, done = [0.0] * args.num_processes, []
    for t in range(args.num_steps):
        for i in range(args.num_processes):
            culm[i] += float(r[t, i])
            if m[t, i] == 0.0:
                done.append(culm[i])
                culm[i] = 0.0
    assert done, "fixture: at least one episode must end inside the rollout"
    assert np.isclose(logs["culm_irl_reward"], np.mean(done), rtol=1e-6)
    assert mod.culm_log_vals["reward"] == pytest.approx(culm, rel=1e-6)
