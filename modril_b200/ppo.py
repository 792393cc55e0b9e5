"""The PPO consumer of the rewards on the device (scope row f4): returns / GAE and advantage normalisation, the
actor-critic policy of ``get_ppo_policy`` and the clipped-PPO minibatch update.

Mirrors ``RolloutStorage.compute_returns`` / ``compute_advantages``
(``deps/rl-toolkit/rlf/storage/rollout_storage.py:178-239``) as two functions over the storage's tensors: the T-step python
loop of small torch ops becomes one ``frl_compute_returns`` launch (bit-identical float32 scan), the advantage
normalisation one ``frl_normalise_advantages`` call.  CUDA tensors only; no PyTorch arithmetic, no CPU fallback.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _native
from .flowril.networks import _FlatVNet, _current_stream, _ptr, as_f32_cuda
from .flowril.pretrain import FusedAdam  # noqa: F401  (the optimiser to pair with ppo_update)


def _f32(x, dev, what):
    if not isinstance(x, torch.Tensor) or x.device.type != "cuda":
        raise RuntimeError(f"{what} must be a CUDA tensor (there is no CPU fallback)")
    if x.device != dev:
        raise RuntimeError(f"{what} is on {x.device}, expected {dev}")
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise TypeError(f"{what} must be a contiguous float32 tensor (it is updated / read in place)")
    return x


def compute_returns(rewards, value_preds, masks, bad_masks, next_value, returns, *, gamma, gae_lambda=0.95, use_gae=True,
                    use_proper_time_limits=False):
    """``storage.compute_returns(next_value)``: rewards [T,N,1], value_preds / returns [T+1,N,D], masks / bad_masks
    [T+1,N,1], next_value [N,D].  Like the reference it writes ``value_preds[-1] = next_value`` (GAE) or
    ``returns[-1] = next_value`` (plain discounting) and fills ``returns[:-1]`` in place; returns ``returns``."""
    dev = returns.device
    T = rewards.shape[0]
    N, D = value_preds.shape[1], value_preds.shape[2]
    if rewards.shape != (T, N, 1) or value_preds.shape != (T + 1, N, D) or returns.shape != (T + 1, N, D) \
            or masks.shape != (T + 1, N, 1) or next_value.shape != (N, D):
        raise ValueError("shape mismatch: rewards [T,N,1], value_preds/returns [T+1,N,D], masks [T+1,N,1], next_value [N,D]")
    bm = None
    if use_proper_time_limits:
        if bad_masks is None or bad_masks.shape != (T + 1, N, 1):
            raise ValueError("use_proper_time_limits needs bad_masks [T+1,N,1]")
        bm = _f32(bad_masks, dev, "bad_masks")
    _native.check(_native.lib().frl_compute_returns(
        _ptr(_f32(rewards, dev, "rewards")), _ptr(_f32(value_preds, dev, "value_preds")), _ptr(_f32(masks, dev, "masks")),
        _ptr(bm), _ptr(_f32(next_value, dev, "next_value")), T, N, D, float(gamma), float(gae_lambda), int(bool(use_gae)),
        _ptr(_f32(returns, dev, "returns")), dev.index or 0, _current_stream(dev)), "frl_compute_returns")
    return returns


def compute_advantages(returns, value_preds, eps: float = 1e-5, return_stats: bool = False):
    """``storage.compute_advantages()``: ``adv = returns[:-1] - value_preds[:-1]`` normalised to zero mean / unit
    (unbiased) std.  Returns a new tensor [T,N,D] (and the device tensor ``[mean, std]`` with ``return_stats``)."""
    dev = returns.device
    _f32(returns, dev, "returns"), _f32(value_preds, dev, "value_preds")
    if returns.shape != value_preds.shape or returns.dim() != 3 or returns.shape[0] < 2:
        raise ValueError("returns and value_preds must both be [T+1,N,D]")
    adv = torch.empty_like(returns[:-1])
    scratch = torch.empty(256, dtype=torch.float64, device=dev)
    stats = torch.empty(2, device=dev)
    _native.check(_native.lib().frl_normalise_advantages(_ptr(returns), _ptr(value_preds), adv.numel(), float(eps), _ptr(adv),
                                                         _ptr(scratch), _ptr(stats), dev.index or 0, _current_stream(dev)),
                  "frl_normalise_advantages")
    return (adv, stats) if return_stats else adv


# ---- actor-critic policy + clipped-PPO update (row f4, second part) ---------------------------------------------------
class _MLPBasic(nn.Module):
    """``MLPBasic`` (deps/rl-toolkit/rlf/rl/model.py:246-296): ``net`` = [Linear, Tanh] * num_layers, orthogonal weights
    with gain sqrt(2) and zero biases (``def_mlp_weight_init``, model.py:22-25).  Parameter container only."""

    def __init__(self, num_inputs, hidden_size, num_layers):
        super().__init__()
        layers, k = [], num_inputs
        for _ in range(num_layers):
            lin = nn.Linear(k, hidden_size)
            nn.init.orthogonal_(lin.weight.data, gain=math.sqrt(2))
            nn.init.constant_(lin.bias.data, 0)
            layers += [lin, nn.Tanh()]
            k = hidden_size
        self.net = nn.Sequential(*layers)


class _DiagGaussian(nn.Module):
    """``DiagGaussian`` (rlf/rl/distributions.py:60-75): ``logstd`` [1,A] zeros, ``fc_mean`` orthogonal (gain 1)."""

    def __init__(self, num_inputs, num_outputs):
        super().__init__()
        self.fc_mean = nn.Linear(num_inputs, num_outputs)
        nn.init.orthogonal_(self.fc_mean.weight.data)
        nn.init.constant_(self.fc_mean.bias.data, 0)
        self.logstd = nn.Parameter(torch.zeros(1, num_outputs))


class DistActorCritic(_FlatVNet):
    """The policy ``get_ppo_policy`` builds for ``--alg flowril`` (deps/baselines/get_policy.py:12-23):
    ``DistActorCritic`` (rlf/policies/actor_critic/dist_actor_critic.py:13-89) with the identity base net, an
    ``MLPBasic`` critic + ``critic_head`` (base_actor_critic.py:37-50) and an ``MLPBasic`` actor + ``DiagGaussian``.
    Same sub-module names and registration order, hence the same ``state_dict`` keys (``critic.net.0.weight`` ...
    ``dist.logstd``, ``dist.fc_mean.weight``).  All arithmetic runs in ``frl_acnet_*`` / ``frl_ppo_loss_grad``."""

    _ABI = ("frl_acnet_create", "frl_acnet_destroy", "frl_acnet_set_params")

    def __init__(self, obs_dim: int, action_dim: int, ppo_hidden_dim: int = 64, ppo_layers: int = 2,
                 deterministic_policy: bool = False):
        super().__init__()
        self.obs_dim, self.action_dim = int(obs_dim), int(action_dim)
        self.ppo_layers = int(ppo_layers)
        self.deterministic_policy = bool(deterministic_policy)
        self.critic = _MLPBasic(obs_dim, ppo_hidden_dim, ppo_layers)
        self.critic_head = nn.Linear(ppo_hidden_dim, 1)                      # rlf/policies/utils.py:16-19
        nn.init.orthogonal_(self.critic_head.weight.data, gain=math.sqrt(2))
        nn.init.constant_(self.critic_head.bias.data, 0)
        self.actor = _MLPBasic(obs_dim, ppo_hidden_dim, ppo_layers)
        self.dist = _DiagGaussian(ppo_hidden_dim, action_dim)
        self._finish_init(obs_dim, action_dim, ppo_hidden_dim)

    def _native_config(self, device_index: int):
        return _native.ACNetConfig(self.obs_dim, self.action_dim, self.hidden, self.ppo_layers, device_index)

    def _obs(self, state):
        state = as_f32_cuda(state, self._flat.device, "state")
        if state.dim() != 2 or state.shape[1] != self.obs_dim:
            raise ValueError(f"expected state [B,{self.obs_dim}], got {tuple(state.shape)}")
        return state

    def forward(self, state, add_state=None, hxs=None, masks=None):
        """-> (action mean [B,A], value [B,1]); the distribution is Normal(mean, exp(dist.logstd))."""
        h = self.native()
        state = self._obs(state)
        B = state.shape[0]
        value = torch.empty(B, 1, device=state.device)
        mean = torch.empty(B, self.action_dim, device=state.device)
        _native.check(_native.lib().frl_acnet_forward(h, _ptr(state), B, _ptr(value), _ptr(mean), _current_stream(state.device)),
                      "frl_acnet_forward")
        return mean, value

    def get_value(self, inputs, add_inputs=None, hxs=None, masks=None):
        """base_actor_critic.py:52-56 (only the critic runs)."""
        h = self.native()
        state = self._obs(inputs)
        value = torch.empty(state.shape[0], 1, device=state.device)
        _native.check(_native.lib().frl_acnet_forward(h, _ptr(state), state.shape[0], _ptr(value), None,
                                                      _current_stream(state.device)), "frl_acnet_forward")
        return value

    def get_action(self, state, add_state=None, hxs=None, masks=None, step_info=None, *, noise=None):
        """dist_actor_critic.py:48-61 -> dict(value [B,1], action [B,A], action_log_probs [B,1], dist_entropy [B]).
        ``noise`` is the N(0,1) draw of ``Normal.sample`` (drawn here when omitted; ignored by a deterministic policy)."""
        h = self.native()
        state = self._obs(state)
        B, dev = state.shape[0], state.device
        if self.deterministic_policy:
            noise = None
        elif noise is None:
            noise = torch.randn(B, self.action_dim, device=dev)
        else:
            noise = as_f32_cuda(noise, dev, "noise")
            if noise.shape != (B, self.action_dim):
                raise ValueError(f"expected noise [B,{self.action_dim}], got {tuple(noise.shape)}")
        value, logp = torch.empty(B, 1, device=dev), torch.empty(B, 1, device=dev)
        action, ent = torch.empty(B, self.action_dim, device=dev), torch.empty(B, device=dev)
        _native.check(_native.lib().frl_acnet_act(h, _ptr(state), _ptr(noise), B, _ptr(value), _ptr(action), _ptr(logp),
                                                  _ptr(ent), _current_stream(dev)), "frl_acnet_act")
        return {"value": value, "action": action, "action_log_probs": logp, "dist_entropy": ent}

    def evaluate_actions(self, state, add_state, hxs, masks, action):
        """dist_actor_critic.py:76-86 -> {'value' [B,1], 'log_prob' [B,1], 'ent' [B]}."""
        h = self.native()
        state = self._obs(state)
        B, dev = state.shape[0], state.device
        action = as_f32_cuda(action, dev, "action")
        if action.shape != (B, self.action_dim):
            raise ValueError(f"expected action [B,{self.action_dim}], got {tuple(action.shape)}")
        value, logp, ent = torch.empty(B, 1, device=dev), torch.empty(B, 1, device=dev), torch.empty(B, device=dev)
        _native.check(_native.lib().frl_acnet_evaluate(h, _ptr(state), _ptr(action), B, _ptr(value), _ptr(logp), _ptr(ent),
                                                       _current_stream(dev)), "frl_acnet_evaluate")
        return {"value": value, "log_prob": logp, "ent": ent}

    def ppo_loss(self, state, action, prev_log_prob, adv, ret, value, *, indices=None, clip_param=0.2, value_loss_coef=0.5,
                 entropy_coef=0.01, grad=True, accumulate=False, mean_divisor: int = 0, grad_scale: float = 1.0):
        """Loss of one PPO.update minibatch (ppo.py:23-49) and, with ``grad=True``, its parameter gradient into
        ``flat_grad()``.  With ``indices`` (int64 [B]) the tensors are the flattened rollout arrays and the minibatch is
        gathered inside the kernels.  Returns the device tensor ``[loss, value_loss, actor_loss, dist_entropy]``."""
        h = self.native()
        dev = self._flat.device
        state = self._obs(state)
        rows = state.shape[0]
        action = as_f32_cuda(action, dev, "action")
        cols = [as_f32_cuda(x, dev, n).reshape(-1) for n, x in (("prev_log_prob", prev_log_prob), ("adv", adv),
                                                                 ("return", ret), ("value", value))]
        if action.shape != (rows, self.action_dim) or any(c.numel() != rows for c in cols):
            raise ValueError("state / action / prev_log_prob / adv / return / value disagree on the number of rows")
        if indices is not None:
            if indices.dtype != torch.int64 or indices.device != dev or not indices.is_contiguous():
                raise TypeError("indices must be a contiguous int64 CUDA tensor")
            B = indices.numel()
        else:
            B = rows
        losses = torch.empty(4, device=dev)
        g = self.flat_grad() if grad else None
        _native.check(_native.lib().frl_ppo_loss_grad(h, _ptr(state), _ptr(action), *(_ptr(c) for c in cols), _ptr(indices), B,
                                                      int(mean_divisor), float(clip_param), float(value_loss_coef),
                                                      float(entropy_coef), float(grad_scale), int(bool(accumulate)),
                                                      _ptr(losses), _ptr(g), _current_stream(dev)), "frl_ppo_loss_grad")
        return losses


def ppo_update(policy: DistActorCritic, opt: FusedAdam, obs, actions, action_log_probs, value_preds, returns, advantages, *,
               num_epochs=4, num_mini_batch=4, clip_param=0.2, value_loss_coef=0.5, entropy_coef=0.01, max_grad_norm=0.5,
               perms=None):
    """``PPO.update`` after the returns / advantages are in place (ppo.py:17-64): ``num_epochs`` passes of
    ``num_mini_batch`` minibatches (``BatchSampler(SubsetRandomSampler(range(T*N)), T*N // num_mini_batch,
    drop_last=True)``, rollout_storage.py:258-272), each one loss + backward + ``clip_grad_norm_(max_grad_norm)`` +
    ``Adam.step`` (``_standard_step``, base_net_algo.py:84-100).  Inputs are the rollout tensors with the time and
    process axes flattened: obs [T*N,S] (``storage.obs[:-1]``), actions [T*N,A], action_log_probs / value_preds[:-1] /
    returns[:-1] / advantages [T*N(,1)].  ``perms``: optional list of ``num_epochs`` int64 CUDA permutations (tests; the
    reference draws them from the global RNG).  No host synchronisation inside the loop; returns the reference's
    ``log_vals`` dict (one device->host copy at the end)."""
    dev = policy.flat_params().device
    rows = obs.shape[0]
    if rows < num_mini_batch:
        raise ValueError(f"PPO requires T*N ({rows}) >= num_mini_batch ({num_mini_batch})")   # rollout_storage.py:262-268
    mb = rows // num_mini_batch
    acc = torch.zeros(4, device=dev)
    n_updates = 0
    for e in range(num_epochs):
        perm = perms[e] if perms is not None else torch.randperm(rows, device=dev)
        for k in range(rows // mb):                                   # drop_last=True
            losses = policy.ppo_loss(obs, actions, action_log_probs, advantages, returns, value_preds,
                                     indices=perm[k * mb:(k + 1) * mb].contiguous(), clip_param=clip_param,
                                     value_loss_coef=value_loss_coef, entropy_coef=entropy_coef)
            opt.step_flat([policy.flat_grad()], max_norm=max_grad_norm if max_grad_norm > 0 else float("inf"))
            acc += losses
            n_updates += 1
    a = acc.tolist()
    n = num_epochs * num_mini_batch                                   # ppo.py:59-64
    return {"value_loss": a[1] / n, "actor_loss": a[2] / n, "dist_entropy": a[3] / n,
            "policy_update_data": float(num_mini_batch * n_updates)}


class PPOUpdateGraph:
    """``ppo_update`` captured ONCE as a CUDA graph and replayed per ``PPO.update``: all ``num_epochs x num_mini_batch``
    minibatches (weight re-pack, forward, loss, backward, clip, Adam: ~40 small launches each at the reference's size) cost
    one graph launch instead of ~600 kernel launches driven from Python.  The rollout tensors are copied into static
    buffers and the sampler's permutations are written into a static index buffer before the replay.

    ``opt`` must be a ``FusedAdam([policy], ..., capturable=True)`` (step count on the device)."""

    def __init__(self, policy: DistActorCritic, opt: FusedAdam, rows: int, *, num_epochs=4, num_mini_batch=4, clip_param=0.2,
                 value_loss_coef=0.5, entropy_coef=0.01, max_grad_norm=0.5):
        if opt._step_dev is None:
            raise ValueError("PPOUpdateGraph needs FusedAdam(..., capturable=True)")
        if rows < num_mini_batch:
            raise ValueError(f"PPO requires T*N ({rows}) >= num_mini_batch ({num_mini_batch})")
        dev = policy.flat_params().device
        self.policy, self.opt, self.rows = policy, opt, int(rows)
        self.num_epochs, self.num_mini_batch = int(num_epochs), int(num_mini_batch)
        self._obs = torch.zeros(rows, policy.obs_dim, device=dev)
        self._act = torch.zeros(rows, policy.action_dim, device=dev)
        self._cols = [torch.zeros(rows, device=dev) for _ in range(4)]      # prev_log_prob, value_preds, returns, adv
        self._perm = torch.arange(rows, device=dev).repeat(num_epochs, 1).contiguous()
        self._acc = torch.zeros(4, device=dev)
        mb = rows // num_mini_batch
        max_norm = max_grad_norm if max_grad_norm > 0 else float("inf")
        kw = dict(clip_param=clip_param, value_loss_coef=value_loss_coef, entropy_coef=entropy_coef)

        def minibatch(e, k):
            losses = policy.ppo_loss(self._obs, self._act, self._cols[0], self._cols[3], self._cols[2], self._cols[1],
                                     indices=self._perm[e, k * mb:(k + 1) * mb], **kw)
            opt.step_flat([policy.flat_grad()], max_norm=max_norm)
            self._acc += losses

        # warm-up outside the capture (workspaces, the optimiser's buffers), then put the state back
        snap = [t.clone() for t in (policy.flat_params(), opt._m[0], opt._v[0], opt._step_dev)]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            minibatch(0, 0)
        torch.cuda.current_stream(dev).wait_stream(side)
        with torch.no_grad():
            for dst, src in zip((policy.flat_params(), opt._m[0], opt._v[0], opt._step_dev), snap):
                dst.copy_(src)
        policy.mark_dirty()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._acc.zero_()
            for e in range(self.num_epochs):
                for k in range(rows // mb):
                    minibatch(e, k)
        # the capture itself does not execute anything, but it left the module's pack key pointing at "packed"
        policy.mark_dirty()

    def __call__(self, obs, actions, action_log_probs, value_preds, returns, advantages, perms=None):
        """Same arguments and return value as ``ppo_update``."""
        with torch.no_grad():
            self._obs.copy_(obs.reshape(self.rows, -1))
            self._act.copy_(actions.reshape(self.rows, -1))
            for dst, src in zip(self._cols, (action_log_probs, value_preds, returns, advantages)):
                dst.copy_(src.reshape(-1))
            for e in range(self.num_epochs):
                if perms is not None:
                    self._perm[e].copy_(perms[e])
                else:
                    torch.randperm(self.rows, device=self._perm.device, out=self._perm[e])
        self._graph.replay()
        self.policy.mark_dirty()
        a = self._acc.tolist()
        n = self.num_epochs * self.num_mini_batch
        n_updates = self.num_epochs * (self.rows // (self.rows // self.num_mini_batch))
        return {"value_loss": a[1] / n, "actor_loss": a[2] / n, "dist_entropy": a[3] / n,
                "policy_update_data": float(self.num_mini_batch * n_updates)}


class PPO:
    """The agent updater of ``FLOWRIL(NestedAlgo([FlowMatchingEstimation(), PPO()]))`` (reference flowril.py:21-23):
    mirror of ``rlf.algos.on_policy.ppo.PPO`` (ppo.py:8-80) with ``OnPolicy._compute_returns`` (on_policy_base.py:23-40)
    and the optimiser handling of ``BaseNetAlgo`` (base_net_algo.py:31-146: Adam(lr, eps), ``--max-grad-norm``,
    ``--linear-lr-decay``) for a ``DistActorCritic`` policy.  ``update(rollouts)`` consumes the rewards the reward module
    wrote into ``rollouts.rewards`` and runs entirely on the device: next value -> returns / GAE scan -> advantage
    normalisation -> ``num_epochs x num_mini_batch`` clipped-PPO minibatches; one device->host copy (the logged means).

    ``rollouts`` needs the tensors of ``RolloutStorage`` (rlf/storage/rollout_storage.py:31-110) on the policy's device:
    ``obs`` [T+1,N,S], ``actions`` [T,N,A], ``action_log_probs`` [T,N,1], ``rewards`` [T,N,1], ``value_preds`` /
    ``returns`` [T+1,N,1], ``masks`` / ``bad_masks`` [T+1,N,1]."""

    def __init__(self):
        self.arg_prefix = ""
        self._graph = None

    def init(self, policy: DistActorCritic, args):
        self.policy, self.args = policy, args
        self._initial_lr = float(args.lr)
        self.opt = FusedAdam([policy], lr=float(args.lr), eps=float(args.eps), capturable=True)
        # total number of updates of the run (base_net_algo.py:48-58); only the lr schedule needs it
        per_update = int(args.num_steps) * int(args.num_processes)
        self.lr_updates = max(1, int(getattr(args, "num_env_steps", per_update)) // per_update)

    def pre_update(self, cur_update):
        if getattr(self.args, "linear_lr_decay", True):              # autils.linear_lr_schedule (rlf/algos/utils.py:76-79)
            lr = self._initial_lr - (self._initial_lr * (cur_update / float(self.lr_updates)))
            for group in self.opt.param_groups:
                group["lr"] = lr
            self._graph = None                                       # the learning rate is a kernel argument of the graph

    def _compute_returns(self, rollouts):
        a = self.args
        next_value = self.policy.get_value(rollouts.obs[-1])         # on_policy_base.py:23-36
        compute_returns(rollouts.rewards, rollouts.value_preds, rollouts.masks, getattr(rollouts, "bad_masks", None),
                        next_value, rollouts.returns, gamma=a.gamma, gae_lambda=a.gae_lambda, use_gae=a.use_gae,
                        use_proper_time_limits=getattr(a, "use_proper_time_limits", False))

    def update(self, rollouts, args=None, beginning=False, t=1, *, perms=None, graph=None):
        """ppo.py:9-64.  ``graph``: replay the minibatch loop as one CUDA graph (default: when the learning rate is
        constant, i.e. ``--linear-lr-decay False``)."""
        a = self.args
        self._compute_returns(rollouts)
        advantages = compute_advantages(rollouts.returns, rollouts.value_preds)
        T, N = rollouts.rewards.shape[:2]
        rows = T * N
        flat = (rollouts.obs[:-1].reshape(rows, -1), rollouts.actions.reshape(rows, -1), rollouts.action_log_probs.reshape(rows),
                rollouts.value_preds[:-1].reshape(rows), rollouts.returns[:-1].reshape(rows), advantages.reshape(rows))
        kw = dict(num_epochs=a.num_epochs, num_mini_batch=a.num_mini_batch, clip_param=a.clip_param,
                  value_loss_coef=a.value_loss_coef, entropy_coef=a.entropy_coef, max_grad_norm=a.max_grad_norm)
        if graph is None:
            graph = not getattr(a, "linear_lr_decay", True)
        if graph:
            if self._graph is None or self._graph.rows != rows:
                self._graph = PPOUpdateGraph(self.policy, self.opt, rows, **kw)
            return self._graph(*flat, perms=perms)
        return ppo_update(self.policy, self.opt, *flat, perms=perms, **kw)

    def get_add_args(self, parser):
        """The flags of OnPolicy / PPO / BaseNetAlgo (on_policy_base.py:13-21, ppo.py:66-80, base_net_algo.py:105-146)."""
        str2bool = lambda v: v.lower() == "true"
        p = self.arg_prefix
        parser.add_argument(f"--{p}num-epochs", type=int, default=4)
        parser.add_argument(f"--{p}num-mini-batch", type=int, default=4)
        parser.add_argument(f"--{p}clip-param", type=float, default=0.2)
        parser.add_argument(f"--{p}entropy-coef", type=float, default=0.01)
        parser.add_argument(f"--{p}value-loss-coef", type=float, default=0.5)
        parser.add_argument(f"--{p}max-grad-norm", type=float, default=0.5)
        parser.add_argument(f"--{p}linear-lr-decay", type=str2bool, default=True)
        parser.add_argument(f"--{p}eps", type=float, default=1e-12)
        parser.add_argument(f"--{p}lr", type=float, default=1e-3)

    # the rest of rlf's BaseAlgo surface that NestedAlgo fans out to its modules (nested_algo.py:25-89); nothing to do here
    def set_env_ref(self, envs):
        self.envs = envs

    def set_get_policy(self, get_policy_fn, policy_args):
        self._get_policy_fn, self._policy_args = get_policy_fn, policy_args

    def first_train(self, log, eval_policy, env_interface):
        pass

    def on_traj_finished(self, traj):
        pass

    def load(self, checkpointer):
        if checkpointer.has_load_key("actor_opt"):                   # base_net_algo.py:71-75
            self.load_resume(checkpointer)

    def get_steps_generator(self, update_iter):
        return range(self.args.num_steps)                            # base_algo.py:62-64

    def get_num_updates(self):                                       # base_algo.py:97-108
        a = self.args
        return 0 if a.num_steps == 0 else int(a.num_env_steps) // a.num_steps // a.num_processes

    def get_completed_update_steps(self, num_updates):
        return num_updates * self.args.num_processes * self.args.num_steps

    def save(self, checkpointer):
        checkpointer.save_key("actor_opt", self.opt.state_dict())    # base_net_algo.py:66-69

    def load_resume(self, checkpointer):
        self.opt.load_state_dict(checkpointer.get_key("actor_opt"))  # base_net_algo.py:60-64
        self._graph = None
