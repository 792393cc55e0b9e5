"""The FLOWRIL reward module (plugin for the rl-toolkit PPO updater) on the B200 kernels.

Mirror of the reference's ``modril/flowril/flowril.py`` as far as its hooks touch the hot path (SURVEY.md section 8:
rows a7-a12): same class names (``FLOWRIL``, ``FlowMatchingEstimation``), hook names (``init``,
``_create_flow_net``, ``_update_reward_func``, ``_get_reward``, ``_compute_flow_reward``, ``_compute_flow_loss``,
``_compute_disc_val``, ``get_add_args``, ``save`` / ``load_resume``), CLI flags / YAML keys (flowril.py:545-579) and
checkpoint keys (``flow_<option>``, ``flow_<option>_opt``; flowril.py:581-597).  Plotting / animation code
(flowril.py:367-543) is out of scope.

When rl-toolkit (``rlf``) is importable the classes derive from its ``BaseIRLAlgo`` / ``NestedAlgo`` and plug into
``deps/baselines/main.py --alg flowril`` unchanged.  rl-toolkit needs gym + mujoco at import time, so for the tests
and for stand-alone use a minimal stand-in base class reproduces the part of ``BaseIRLAlgo.update``
(deps/rl-toolkit/rlf/algos/il/base_irl.py:37-72) that drives the hooks.

Differences in mechanism, not in results:
  * ``_get_reward(step, ...)`` is served from ONE fused launch over the whole rollout (T*N rows, both roles) made at
    the first step of an update; the log-density is a pure function of (s, a, weights, probe) and the weights do not
    change between the steps of one update, so the rows are bit-identical to per-step scoring.  The return
    normalisation (flowril.py:221-228: discounted-return recursion, float32 RunningMeanStd, scaling) is a sequential
    scan over the steps; it runs as ONE more launch (``frl_normalise_rewards``) with its state on the device, so an
    update does no device->host copy for the rewards at all (the reference does one per step).
  * ``_update_reward_func`` runs the all-native route (fused CFM gradient kernels + clip+Adam kernel) on the flat
    parameter buffer.
  * Reference quirk kept on purpose: ``_update_reward_func`` passes
    ``(expert_batch, agent_batch)`` to ``_compute_flow_loss(agent_batch, expert_batch, ...)`` (flowril.py:314 vs :232),
    so role "expert" is trained on the ROLLOUT batch and role "agent" on the DEMONSTRATION batch.
  * The anti-divergence / Jacobian-stability regularisers (flowril.py:255-290; ON by default) come out of
    ``frl_reg_loss_grad`` as loss values plus hand-derived parameter gradients instead of autograd's double backward;
    their weights (fixed, warm-up ``L_E / (L_reg + eps) * 1e-3``, or GradNorm2D's beta / gamma) are applied to the
    gradients on the device.  Reference quirk kept (flowril.py:272): ``gradnorm.step_count`` only advances inside
    ``GradNorm2D.forward``, which is only called from the branch taken when ``step_count > warmup_steps`` -- so with
    the reference's control flow the warm-up weighting is what always runs; the GradNorm branch is implemented and
    tested but, exactly like in the reference, not reachable from ``_compute_flow_loss`` unless ``step_count`` is
    advanced externally.
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np
import torch

from .networks import GradNorm2D
from .pretrain import (CoupledResidualFM, FlowMatching, FusedAdam, _rademacher, score_pair, tensor_grad_norm_sum,
                       train_step)

try:  # the real plugin API, when rl-toolkit and its simulators are installed
    import rlf.rl.utils as rutils
    from rlf.algos.il.base_irl import BaseIRLAlgo
    from rlf.algos.nested_algo import NestedAlgo
    from rlf.algos.on_policy.ppo import PPO
    from rlf.baselines.common.running_mean_std import RunningMeanStd

    HAVE_RLF = True
except Exception:  # pragma: no cover - exercised in this repo's tests
    HAVE_RLF = False

    class _Utils:
        @staticmethod
        def get_obs_shape(space):
            return tuple(space.shape)

        @staticmethod
        def get_ac_dim(space):
            return int(space.shape[0])

        @staticmethod
        def get_ac_repr(space, action):
            return action  # identity for Box spaces (rlf/rl/utils.py:392-404)

        @staticmethod
        def get_def_obs(obs):
            return obs["observation"] if isinstance(obs, dict) and "observation" in obs else obs

    rutils = _Utils()

    class BaseIRLAlgo(object):
        """The slice of rl-toolkit's BaseIRLAlgo / BaseILAlgo this module relies on
        (base_irl.py:37-72, base_il.py:63-104): expert loader + the update loop that calls the two hooks."""

        def __init__(self):
            self.expert_train_loader = None

        def init(self, policy, args):
            self.policy, self.args = policy, args
            loader = getattr(args, "expert_loader", None)
            if loader is None and getattr(args, "expert_dataset", None) is not None:
                ds = args.expert_dataset
                n = ds["obs"].shape[0]

                class _DS(torch.utils.data.Dataset):
                    def __len__(self):
                        return n

                    def __getitem__(self, i):
                        return {"state": ds["obs"][i], "actions": ds["actions"][i]}

                loader = torch.utils.data.DataLoader(_DS(), batch_size=args.traj_batch_size, shuffle=True, drop_last=True)
            self.expert_train_loader = loader

        def get_env_ob_filt(self):
            return None

        def get_env_settings(self, args):
            return type("Settings", (), {})()

        def get_add_args(self, parser):
            pass

        def _adjust_action(self, x):
            return x

        def update(self, storage, gradient_clip=False, t=1):
            """base_irl.py:37-72: wipe the environment rewards, train the reward model (the hook is first tried with the
            caller's extra arguments, then without), relabel every rollout step through ``_get_reward``, and keep the
            per-process running sums that become ``culm_irl_<key>`` (mean over the last ``log_smooth_len`` finished
            episodes) in the log."""
            T, N = self.args.num_steps, self.args.num_processes
            if not hasattr(self, "ep_log_vals"):
                from collections import deque

                self.ep_log_vals = defaultdict(lambda: deque(maxlen=getattr(self.args, "log_smooth_len", 10)))
                self.culm_log_vals = defaultdict(lambda: [0.0] * N)
            for step in range(T):
                storage.rewards[step] = 0
            try:
                log_vals = self._update_reward_func(storage, gradient_clip, t)
            except TypeError:
                log_vals = self._update_reward_func(storage)
            per_step = defaultdict(list)
            for step in range(T):
                rewards, ep_vals = self._get_reward(step, storage, {})
                storage.rewards[step] = rewards
                for k, v in dict(ep_vals, reward=rewards).items():
                    per_step[k].append(v.detach().reshape(-1))
            # the logging of base_irl.py:57-67 from ONE device->host copy per key (the reference calls .item() per
            # process and step, and compares the mask on the device each time)
            ended = (storage.masks[:T].reshape(T, -1) == 0.0).cpu().numpy()
            host = {k: torch.stack(v).cpu().numpy() for k, v in per_step.items()}
            for step in range(T):
                for i in range(N):
                    for k, vals in host.items():
                        self.culm_log_vals[k][i] += float(vals[step, i])
                    if ended[step, i]:
                        for k in host:
                            self.ep_log_vals[k].append(self.culm_log_vals[k][i])
                            self.culm_log_vals[k][i] = 0.0
            for k, vals in self.ep_log_vals.items():
                log_vals[f"culm_irl_{k}"] = np.mean(vals)
            return log_vals

        def save(self, checkpointer):
            pass

        def load_resume(self, checkpointer):
            pass

    NestedAlgo = None
    PPO = None


def _canonical_device(device):
    """torch.device('cuda') -> torch.device('cuda', current index): tensors report an indexed device, and the native
    layer compares devices exactly."""
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


class DeviceRunningMeanStd(object):
    """The float32 running moments of rlf's RunningMeanStd (running_mean_std.py:3-35) kept ON THE DEVICE and advanced
    by ``frl_normalise_rewards`` together with the discounted-return recursion (reference flowril.py:221-228), so the
    reward hook needs no device->host copy per rollout step.  ``mean`` / ``var`` / ``count`` read the state back
    (one sync) with the shapes the numpy object has after its first update (``var[0]`` works)."""

    def __init__(self, device, epsilon=1e-4):
        self.device = _canonical_device(device)
        self.mean_var = torch.tensor([0.0, 1.0], device=self.device)
        self.count_t = torch.tensor([epsilon], dtype=torch.float64, device=self.device)

    @property
    def mean(self):
        return self.mean_var[:1].cpu().numpy()

    @property
    def var(self):
        return self.mean_var[1:].cpu().numpy()

    @property
    def count(self):
        return float(self.count_t.item())

    def normalise(self, reward, masks, gamma, returns):
        """reward / masks [T, N, 1] device tensors; returns: [N, 1] carried state or None (first call).
        Returns (normalised reward [T, N, 1], new returns [N, 1])."""
        from .. import _native
        from .networks import _current_stream, _ptr, as_f32_cuda

        T, N = reward.shape[0], reward.shape[1]
        r = as_f32_cuda(reward, self.device, "reward").reshape(T, N)
        m = as_f32_cuda(masks, self.device, "masks").reshape(T, N)
        valid = returns is not None
        ret = returns.detach().float().reshape(N).clone() if valid else torch.empty(N, device=self.device)
        out = torch.empty(T, N, device=self.device)
        _native.check(_native.lib().frl_normalise_rewards(_ptr(r), _ptr(m), T, N, float(gamma), _ptr(ret),
                                                          1 if valid else 0, _ptr(self.mean_var), _ptr(self.count_t),
                                                          _ptr(out), self.device.index or 0,
                                                          _current_stream(self.device)), "frl_normalise_rewards")
        return out.reshape(T, N, 1), ret.reshape(N, 1)


def str2bool(v):
    return v.lower() == "true"


class FlowMatchingEstimation(BaseIRLAlgo):
    def __init__(self):
        super().__init__()
        self.step = 0
        self._rollout_cache = None
        self._norm_cache = None
        self._serve = None

    # ---- construction (reference flowril.py:34-121) -------------------------------------------------------------
    def _create_flow_net(self):
        ob_shape = rutils.get_obs_shape(self.policy.obs_space)
        self.state_dim = ob_shape[0]
        self.action_dim = rutils.get_ac_dim(self.action_space)
        dev = self.args.device
        lr = self.args.disc_lr
        if self.args.option == "scrf":
            # NB: like the reference (flowril.py:42-46) scrf does not pass --flow-depth (class default 4)
            self.flow_net = CoupledResidualFM(state_dim=self.state_dim, action_dim=self.action_dim,
                                              hidden_dim=self.args.hidden_dim).to(dev)
            if self.args.flow_path:
                self.flow_net.load_state_dict(torch.load(self.args.flow_path, map_location=dev))
                for p in self.flow_net.net.shared.parameters():   # freeze the common field, train the residual
                    p.requires_grad = False
                for p in self.flow_net.net.head_c.parameters():
                    p.requires_grad = False
                for p in self.flow_net.net.head_r.parameters():
                    p.requires_grad = True
            self.opt = FusedAdam([self.flow_net.net], lr=lr)
            self.params = list(self.flow_net.parameters())
            if self.args.params_autotune:
                self.gradnorm = GradNorm2D(beta_init=self.args.loss_anti, gamma_init=self.args.loss_stable, alpha=0.1,
                                           warmup_steps=100, update_freq=10, ema_decay=0.9, beta_min=1e-8,
                                           beta_max=1e1, gamma_min=1e-12, gamma_max=1e-2,
                                           enable_beta=self.args.enable_loss_anti,
                                           enable_gamma=self.args.enable_loss_stable).to(dev)
            if not self.args.enable_loss_anti:
                self.args.loss_anti = 0
            if not self.args.enable_loss_stable:
                self.args.loss_stable = 0
            # device copies of the two weights (2 floats each: value + scratch): `use_anti = args.loss_anti > 0`
            # (flowril.py:255-256) is evaluated on the device as a gate, so no host sync per batch
            self._w_anti_init, self._w_stable_init = float(self.args.loss_anti), float(self.args.loss_stable)
            self._w_anti = torch.tensor([self._w_anti_init, 0.0], device=dev)
            self._w_stable = torch.tensor([self._w_stable_init, 0.0], device=dev)
        else:
            mk = lambda: FlowMatching(state_dim=self.state_dim, action_dim=self.action_dim,
                                      hidden_dim=self.args.hidden_dim, depth=self.args.flow_depth).to(dev)
            self.flow_net_e, self.flow_net_a = mk(), mk()
            if self.args.flow_path and not self.args.finetune_ve:
                self.flow_net_e.load_state_dict(torch.load(self.args.flow_path, map_location=dev))
                for p in self.flow_net_e.parameters():
                    p.requires_grad = False
                self.opt = FusedAdam([self.flow_net_a.net], lr=lr)
            else:
                self.opt = FusedAdam([self.flow_net_e.net, self.flow_net_a.net], lr=lr)

    def init(self, policy, args):
        args.device = _canonical_device(args.device)
        super().init(policy, args)
        self.policy = policy
        self.args = args
        self.action_space = self.policy.action_space
        self._create_flow_net()
        self.returns = None
        self.ret_rms = DeviceRunningMeanStd(self.args.device)
        self.eps = 1e-8
        self.global_scale = 1e-3
        self.precision = getattr(args, "flowril_precision", None)

    def _get_sampler(self, storage):
        agent_experience = storage.get_generator(None, mini_batch_size=self.expert_train_loader.batch_size)
        return self.expert_train_loader, agent_experience

    def _trans_batches(self, expert_batch, agent_batch):
        return expert_batch, agent_batch

    def get_env_settings(self, args):
        settings = super().get_env_settings(args)
        if not args.flowril_state_norm:
            settings.ret_raw_obs = True
        return settings

    def _norm_expert_state(self, state, obsfilt):
        if not self.args.flowril_state_norm or obsfilt is None:
            return state   # (without a filter the reference's numpy round trip, flowril.py:157-161, is the identity)
        state = obsfilt(state.cpu().numpy(), update=False)
        return torch.tensor(state).to(self.args.device)

    def _trans_agent_state(self, state, other_state=None):
        if not self.args.flowril_state_norm:
            if isinstance(state, dict):
                return state["raw_obs"] if other_state is None else other_state["raw_obs"]
            return state
        return rutils.get_def_obs(state)

    # ---- scoring (reference flowril.py:175-230, :352-365) ---------------------------------------------------------
    def _nets(self):
        if self.args.option == "scrf":
            return self.flow_net.net, self.flow_net.net
        return self.flow_net_e.net, self.flow_net_a.net

    def _probe(self, n_rows, n_steps=32):
        """The probes the reference would draw for a batch of n_rows (scrf: shape-deterministic `_rademacher`,
        pretrain.py:51-52; 2fs: global-RNG `randint_like` per step and per net, pretrain.py:90 -- expert net first)."""
        dev = self.args.device
        if self.args.option == "scrf":
            return _rademacher((n_rows, self.action_dim), device=dev)
        like = torch.empty(n_rows, self.action_dim, device=dev)
        return torch.stack([torch.randint_like(like, 0, 2).float().mul_(2).sub_(1) for _ in range(n_steps)])

    def _score(self, state, action, reward_type):
        ne, na = self._nets()
        if self.args.option == "scrf":
            probe = self._probe(state.shape[0])
            return score_pair(ne, na, state, action, probe, n_steps=32, reward_type=reward_type,
                              precision=self.precision)
        # 2fs: each net draws its own per-step probes from the global RNG, expert first (flowril.py:187-189)
        x = torch.cat([state, action], dim=-1)
        le = self.flow_net_e.log_prob(x, precision=self.precision)
        la = self.flow_net_a.log_prob(x, precision=self.precision)
        return le, la, None

    def _reward_from(self, le, la, rw):
        if rw is None:   # 2fs: the two nets were integrated by separate launches
            rw = self._reward_native(le, la, self.args.reward_type)
        return rw.unsqueeze(-1)

    def _batched(self):
        return getattr(self.args, "flowril_batched_scoring", True)

    # ---- 2fs: the per-step probes of a whole rollout in the reference's RNG order, one launch ------------------------
    _replay_ok = {}   # device index -> bool: the Philox replay reproduced real randint_like calls on this device

    def _draw_loop(self, T, N, n_steps):
        """The reference's own sequence of draws (flowril.py:186-189 per rollout step: expert net's n_steps probes, then
        the agent net's), T x 2 x n_steps small launches."""
        like = torch.empty(N, self.action_dim, device=self.args.device)
        pe = torch.empty(n_steps, T, N * self.action_dim, device=self.args.device)
        pa = torch.empty_like(pe)
        for t in range(T):
            for buf in (pe, pa):
                for k in range(n_steps):
                    buf[k, t] = torch.randint_like(like, 0, 2).float().mul_(2).sub_(1).reshape(-1)
        return pe, pa

    def _draw_replay(self, T, N, n_steps):
        from .. import _native
        from .networks import _current_stream, _ptr

        dev = self.args.device
        gen = torch.cuda.default_generators[dev.index]
        seed, offset = gen.initial_seed(), gen.get_offset()
        numel = N * self.action_dim
        pe = torch.empty(n_steps, T, numel, device=dev)
        pa = torch.empty_like(pe)
        _native.check(_native.lib().frl_rademacher_calls(seed, offset, T, 2, n_steps, numel, _ptr(pe), _ptr(pa), dev.index,
                                                         _current_stream(dev)), "frl_rademacher_calls")
        gen.set_offset(offset + 4 * T * 2 * n_steps)
        return pe, pa

    def _rollout_probes_2fs(self, T, N, n_steps=32):
        """([n_steps, T*N, a], same) for the expert and the agent net: the values -- and the generator state afterwards --
        the reference's T x 2 x n_steps ``randint_like`` calls would produce, from ONE launch (csrc/probe.cu).  The replay
        is checked against real draws once per device and process; if torch's generator kernel ever changes, the loop of
        real draws is used instead."""
        dev = self.args.device
        ok = FlowMatchingEstimation._replay_ok.get(dev.index)
        if ok is None:
            gen = torch.cuda.default_generators[dev.index]
            state = gen.get_state()
            want = self._draw_loop(2, N, 3)
            end = gen.get_offset()
            gen.set_state(state)
            got = self._draw_replay(2, N, 3)
            ok = gen.get_offset() == end and all(torch.equal(w, g) for w, g in zip(want, got))
            gen.set_state(state)
            FlowMatchingEstimation._replay_ok[dev.index] = ok
        pe, pa = (self._draw_replay if ok else self._draw_loop)(T, N, n_steps)
        shape = (n_steps, T * N, self.action_dim)
        return pe.reshape(shape), pa.reshape(shape)

    def _reward_native(self, le, la, reward_type):
        from .. import _native
        from .networks import _current_stream, _ptr

        out = torch.empty_like(le)
        _native.check(_native.lib().frl_reward(_ptr(le), _ptr(la), _ptr(out), le.numel(), _native.REWARD_TYPES[reward_type],
                                               le.device.index, _current_stream(le.device)), "frl_reward")
        return out

    def _compute_flow_reward(self, storage, step, add_info):
        if self.args.reward_type not in ("raw", "norm", "exp", "clip", "smooth"):
            raise ValueError(f"Unrecognized reward type {self.args.reward_type}")
        if not self._batched():
            state = self._trans_agent_state(storage.get_obs(step)).to(self.args.device)
            action = rutils.get_ac_repr(self.action_space, storage.actions[step].to(self.args.device))
            return self._reward_from(*self._score(state, action, self.args.reward_type))
        if step == 0 or self._rollout_cache is None:
            T = self.args.num_steps
            states = torch.stack([self._trans_agent_state(storage.get_obs(t)).to(self.args.device) for t in range(T)])
            actions = torch.stack([rutils.get_ac_repr(self.action_space, storage.actions[t].to(self.args.device))
                                   for t in range(T)])
            N = states.shape[1]
            ne, na = self._nets()
            if self.args.option == "scrf":
                probe = self._probe(N).repeat(T, 1)  # the same [N, a] probe at every rollout step, like the reference
                le, la, rw = score_pair(ne, na, states.reshape(T * N, -1), actions.reshape(T * N, -1), probe, n_steps=32,
                                        reward_type=self.args.reward_type, precision=self.precision)
            else:
                # 2fs: every net integrates with its own per-step probes (drawn in the reference's order), two launches
                # over all T*N rows + the reward transform
                pe, pa = self._rollout_probes_2fs(T, N)
                x = torch.cat([states.reshape(T * N, -1), actions.reshape(T * N, -1)], dim=-1)
                le = self.flow_net_e.log_prob(x, probe=pe, precision=self.precision)
                la = self.flow_net_a.log_prob(x, probe=pa, precision=self.precision)
                rw = self._reward_native(le, la, self.args.reward_type)
            self._rollout_cache = rw.reshape(T, N, 1)
            self._norm_cache = None
            self._serve = None
        return self._rollout_cache[step]

    def _get_reward(self, step, storage, add_info):
        with torch.no_grad():
            for m in ((self.flow_net,) if self.args.option == "scrf" else (self.flow_net_e, self.flow_net_a)):
                m.eval()
            reward = self._compute_flow_reward(storage, step, add_info)
            if not self.args.flowril_reward_norm:
                if self._batched():
                    if step == 0 or self._serve is None:
                        self._serve = self._host_block(self._rollout_cache)
                    return self._serve[step], {}
                return reward, {}
            # discounted-return recursion + running moments + scaling (reference flowril.py:221-228) on the device
            if self._batched():
                if self._norm_cache is None:
                    T = self.args.num_steps
                    masks = torch.stack([storage.masks[t].to(self.args.device) for t in range(T)])
                    self._norm_cache, self.returns = self.ret_rms.normalise(self._rollout_cache, masks,
                                                                            self.args.gamma, self.returns)
                    self._serve = self._host_block(self._norm_cache)
                return self._serve[step], {}
            out, self.returns = self.ret_rms.normalise(reward.unsqueeze(0), storage.masks[step].unsqueeze(0),
                                                       self.args.gamma, self.returns)
            return out[0], {}

    def _host_block(self, block):
        """Under rl-toolkit's own update loop (base_irl.py:57-67) every returned reward is read element by element with
        ``.item()`` -- T x N device->host syncs per update if the rows live on the device.  ``--flowril-host-rewards``
        (default: on when ``rlf`` drives the module) brings the whole [T, N, 1] block to the host ONCE; the per-step
        rows handed out are then host tensors (``storage.rewards[step] = rewards`` copies them back, no sync)."""
        host = getattr(self.args, "flowril_host_rewards", None)
        if host is None:
            host = HAVE_RLF
        return block.cpu() if host else block

    def _compute_disc_val(self, state, action):
        state, action = state.to(self.args.device), action.to(self.args.device)
        le, la, _ = self._score(state, action, "raw")
        return (le - la).detach()

    # ---- training (reference flowril.py:232-350) ---------------------------------------------------------------------
    def _compute_flow_loss(self, agent_batch, expert_batch, obsfilt):
        """Builds the two (s, a) batches exactly like the reference (incl. its argument-order quirk, see the module
        docstring) and returns the CFM terms for ``train_step``: [(net, sign, s, a0, rate), ...]."""
        expert_states = self._norm_expert_state(expert_batch["state"], obsfilt)
        expert_actions = self._adjust_action(expert_batch["action"].to(self.args.device))
        expert_actions = rutils.get_ac_repr(self.action_space, expert_actions)
        agent_states = self._trans_agent_state(agent_batch["state"],
                                               agent_batch["other_state"] if "other_state" in agent_batch else None)
        agent_actions = rutils.get_ac_repr(self.action_space, agent_batch["actions"].to(self.args.device))
        s_E, a_E = expert_states.to(self.args.device).float(), expert_actions.float()
        s_A, a_A = agent_states.to(self.args.device).float(), agent_actions.float()
        ne, na = self._nets()
        only_a = bool(self.args.flow_path) and not self.args.finetune_ve and self.args.option == "2fs"
        terms = [dict(net=ne, sign=1.0, s=s_E, a0=a_E, rate=0.0 if only_a else self.args.expert_loss_rate),
                 dict(net=na, sign=-1.0, s=s_A, a0=a_A, rate=1.0 if only_a else self.args.agent_loss_rate)]
        return terms

    def _reg_spec(self, terms):
        """The regulariser part of reference flowril.py:255-290 for the (already drawn) scrf terms: mixed batch,
        t_mix ~ U[0,1) (drawn AFTER the two fm_loss draws, like the reference), the shape-deterministic probe, and
        how the two weights are formed.  Returns the ``reg`` dict of ``train_step`` or None."""
        if self.args.option != "scrf":
            return None
        anti_on = self.args.enable_loss_anti if self.args.params_autotune else True
        stable_on = self.args.enable_loss_stable if self.args.params_autotune else True
        # a weight that started at 0 (term disabled) can never turn on: skip the work altogether
        anti_on = anti_on and float(self._w_anti_init) > 0
        stable_on = stable_on and float(self._w_stable_init) > 0
        if not (anti_on or stable_on):
            return None
        s_E, a_E, s_A, a_A = terms[0]["s"], terms[0]["a0_raw"], terms[1]["s"], terms[1]["a0_raw"]
        mix_s = torch.cat([s_E, s_A], dim=0)
        mix_a = torch.cat([a_E, a_A], dim=0)
        t_mix = torch.rand(mix_s.size(0), 1, device=mix_s.device)
        reg = dict(net=self.flow_net.net, s=mix_s, a=mix_a, t=t_mix,
                   probe=_rademacher((mix_a.size(0), self.action_dim), device=mix_a.device))
        if not self.args.params_autotune:
            mode = lambda w: float(w)                       # fixed weights (flowril.py:266-270)
            reg["anti"] = mode(self._w_anti_init) if anti_on else None
            reg["stable"] = mode(self._w_stable_init) if stable_on else None
        elif self.gradnorm.steps <= self.gradnorm.warmup_steps:   # host mirror: no sync
            warm = ("warmup", self.global_scale, self.eps)  # w = L_E / (L_reg + eps) * global_scale (flowril.py:272-284)
            reg["anti"] = warm if anti_on else None
            reg["stable"] = warm if stable_on else None
        else:
            reg["gradnorm"] = True                          # beta / gamma from GradNorm2D (flowril.py:286-290)
            reg["anti"] = self.gradnorm.log_beta.detach().exp().reshape(1) if anti_on else None
            reg["stable"] = self.gradnorm.log_gamma.detach().exp().reshape(1) if stable_on else None
        reg["gate_anti"], reg["gate_stable"] = self._w_anti, self._w_stable
        return reg

    def _after_reg(self, reg):
        """Bookkeeping the reference does after forming the regularised loss: `args.loss_anti/_stable` take the
        weights used (device scalars here), and past warm-up GradNorm2D accumulates / updates (flowril.py:278-290)."""
        out = reg.get("out") or {}
        if reg.get("anti") is not None:
            self.args.loss_anti = self._w_anti[0]
        if reg.get("stable") is not None:
            self.args.loss_stable = self._w_stable[0]
        if reg.get("gradnorm"):
            zero = torch.zeros((), device=self.args.device)
            la, ls = out.get("loss_anti", zero), out.get("loss_stable", zero)
            self.gradnorm(la, ls)
            net = self.flow_net.net
            ga = tensor_grad_norm_sum(net, out["grad_anti"]) if "grad_anti" in out else 0.0
            gs = tensor_grad_norm_sum(net, out["grad_stable"]) if "grad_stable" in out else 0.0
            self.gradnorm.update(la, ls, None, g_anti=ga, g_jpen=gs)

    def _draw(self, term):
        """t / noise (and 2fs dequantisation) in the reference's draw order (pretrain.py:98-101, :182-183)."""
        a0, dev = term["a0"], self.args.device
        B = a0.shape[0]
        eps = self.flow_net.eps if self.args.option == "scrf" else self.flow_net_e.eps
        if self.args.option != "scrf":
            a0 = a0 + torch.randn_like(a0) * 0.02
        t = torch.rand(B, 1, device=dev) * (1 - 2 * eps) + eps
        # the reference draws the noise from the CPU generator (Normal(0., 1.).sample, pretrain.py:99 / :183) and moves it
        # over; same draw here, but through pinned memory with an asynchronous copy: no stream synchronisation per batch
        noise = torch.distributions.Normal(0., 1.).sample((B, self.action_dim)).pin_memory().to(dev, non_blocking=True)
        return dict(term, a0=a0, a0_raw=term["a0"], t=t, noise=noise)

    def _update_reward_func(self, storage):
        for m in ((self.flow_net,) if self.args.option == "scrf" else (self.flow_net_e, self.flow_net_a)):
            m.train()
        self._rollout_cache = None
        log_vals = defaultdict(float)
        obsfilt = self.get_env_ob_filt()
        expert_sampler, agent_sampler = self._get_sampler(storage)
        if agent_sampler is None:
            return {}
        count = 0
        sums = None
        for _ in range(self.args.n_flowril_epochs):
            for expert_batch, agent_batch in zip(expert_sampler, agent_sampler):
                expert_batch, agent_batch = self._trans_batches(expert_batch, agent_batch)
                terms = [self._draw(t) for t in self._compute_flow_loss(expert_batch, agent_batch, obsfilt)]  # sic
                reg = self._reg_spec(terms)
                losses = train_step(terms, self.opt, max_norm=self.args.disc_grad_pen, reg=reg)
                if reg is not None:
                    self._after_reg(reg)
                stacked = torch.stack(losses)
                sums = stacked if sums is None else sums + stacked   # no host sync inside the loop
                self.step += 1
                count += 1
        if sums is not None:
            e_loss, a_loss = (float(x) for x in sums.cpu())
            log_vals["flow_loss"] = e_loss / max(count, 1)
            log_vals["agent_loss"] = a_loss / max(count, 1)
        return log_vals

    # ---- flags / checkpoints (reference flowril.py:545-597) ---------------------------------------------------------------
    def get_add_args(self, parser):
        super().get_add_args(parser)
        parser.add_argument('--n-flowril-epochs', type=int, default=1)
        parser.add_argument('--flowril-state-norm', type=str2bool, default=True)
        parser.add_argument('--flowril-reward-norm', type=str2bool, default=True)
        parser.add_argument('--flow-path', type=str, default=None)
        parser.add_argument('--flow-depth', type=int, default=4)
        parser.add_argument('--disc-lr', type=float, default=1e-4)
        parser.add_argument('--disc-grad-pen', type=float, default=1.0)
        parser.add_argument('--expert-loss-rate', type=float, default=1.0)
        parser.add_argument('--agent-loss-rate', type=float, default=1.0)
        parser.add_argument('--option', type=str, default='scrf', choices=['scrf', '2fs'])
        parser.add_argument('--hidden-dim', type=int, default=256)
        parser.add_argument('--enable-loss-anti', type=str2bool, default=True)
        parser.add_argument('--enable-loss-stable', type=str2bool, default=True)
        parser.add_argument('--params_autotune', type=str2bool, default=True)
        parser.add_argument('--loss-anti', type=float, default=1e-4)
        parser.add_argument('--loss-stable', type=float, default=1e-6)
        parser.add_argument('--flow-loss-rate', type=float, default=0.2)
        parser.add_argument('--finetune-ve', type=str2bool, default=False)
        parser.add_argument('--reward-type', type=str, default='raw')
        parser.add_argument('--plot-during-train', type=bool, default=True)
        parser.add_argument('--save-animation', type=bool, default=True)
        parser.add_argument('--save-interval', type=int, default=10)
        parser.add_argument('--animation-type', type=str, default='gif')
        parser.add_argument('--animation-fps', type=int, default=20)
        # B200-only knobs (not in the reference)
        parser.add_argument('--flowril-precision', type=str, default=None, choices=[None, 'fp16x3', 'fp32', 'fp16', 'bf16'])
        parser.add_argument('--flowril-batched-scoring', type=str2bool, default=True)
        parser.add_argument('--flowril-host-rewards', type=str2bool, default=None)

    def load_resume(self, checkpointer):
        super().load_resume(checkpointer)
        self.opt.load_state_dict(checkpointer.get_key(f'flow_{self.args.option}_opt'))
        if self.args.option == "scrf":
            self.flow_net.load_state_dict(checkpointer.get_key(f'flow_{self.args.option}'))
        else:
            self.flow_net_e.load_state_dict(checkpointer.get_key(f'flow_{self.args.option}_e'))
            self.flow_net_a.load_state_dict(checkpointer.get_key(f'flow_{self.args.option}_a'))

    def save(self, checkpointer):
        super().save(checkpointer)
        checkpointer.save_key(f'flow_{self.args.option}_opt', self.opt.state_dict())
        if self.args.option == "scrf":
            checkpointer.save_key(f'flow_{self.args.option}', self.flow_net.state_dict())
        else:
            # the reference LOADS here instead of saving (flowril.py:596-597); saving is what a
            # resume needs, so this deliberately departs from the bug
            checkpointer.save_key(f'flow_{self.args.option}_e', self.flow_net_e.state_dict())
            checkpointer.save_key(f'flow_{self.args.option}_a', self.flow_net_a.state_dict())


if HAVE_RLF:
    class FLOWRIL(NestedAlgo):
        def __init__(self, agent_updater=None):
            super().__init__([FlowMatchingEstimation(), agent_updater or PPO()], 1)
else:
    class FLOWRIL(object):
        """Stand-alone stand-in for ``NestedAlgo([FlowMatchingEstimation(), PPO()], 1)`` (reference flowril.py:21-23,
        deps/rl-toolkit/rlf/algos/nested_algo.py:6-106) when rl-toolkit is not importable: the reward module followed by
        the device PPO updater of ``modril_b200.ppo``.  ``policy`` must be a ``modril_b200.ppo.DistActorCritic`` carrying
        ``obs_space`` / ``action_space`` (what rl-toolkit's ``BasePolicy.init`` would have stored)."""

        def __init__(self, agent_updater=None):
            if agent_updater is None:
                from ..ppo import PPO as DevicePPO

                agent_updater = DevicePPO()
            self.modules = [FlowMatchingEstimation(), agent_updater]
            self.designated_rl_idx = 1

        def init(self, policy, args):
            for module in self.modules:
                module.init(policy, args)

        def pre_update(self, cur_update):
            for module in self.modules:
                if hasattr(module, "pre_update"):
                    module.pre_update(cur_update)

        def get_add_args(self, parser):
            for module in self.modules:
                module.get_add_args(parser)

        def update(self, storage, args=None, beginning=False, t=1):
            log_vals = {}                                            # nested_algo.py:91-106
            mods = self.modules[:1] if getattr(args, "freeze_policy", False) else self.modules
            for module in mods:
                log_vals = {**log_vals, **module.update(storage, beginning, t)}
            return log_vals

        def save(self, checkpointer):
            for module in self.modules:
                module.save(checkpointer)

        def load_resume(self, checkpointer):
            for module in self.modules:
                module.load_resume(checkpointer)
