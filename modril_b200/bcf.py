"""BCF: behaviour cloning with a flow-matching actor, backed by the sm_100a kernels (scope row f3).

Mirrors the three pieces of the reference's rl-toolkit that make up this path, name for name:

  * ``VectorNetwork``   -- ``deps/rl-toolkit/rlf/rl/model.py:391-432`` (same constructor arguments, sub-module names and
    ``state_dict`` keys: ``time_embed.{0,2}``, ``state_embed.0``, ``action_embed.0``, ``velocity_head.{0,2}``);
  * ``FlowPolicy.get_action``'s Euler sampler -- ``rlf/policies/flow_policy.py:49-63`` (``sample_actions``);
  * ``BehaviorCloneFlowMatching._bc_step`` -- ``rlf/algos/il/bcf.py:99-128`` + ``_standard_step``
    (``rlf/algos/base_net_algo.py:84-100``) (``bcf_loss`` / ``bc_step``).

No PyTorch arithmetic and no CPU fallback: the module holds the weights (one flat fp32 buffer, like the FLOWRIL nets)
and calls ``frl_vnet_forward`` / ``frl_vnet_sample`` / ``frl_bcf_loss_grad`` / ``frl_clip_adam``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _native
from .flowril.networks import _FlatVNet, _current_stream, _ptr, as_f32_cuda
from .flowril.pretrain import FusedAdam  # noqa: F401  (the optimiser to pair with bc_step)


class VectorNetwork(_FlatVNet):
    """v(x_t, s, t): state / time / action embeddings concatenated into a two-layer velocity head (model.py:391-432)."""

    _ABI = ("frl_vnet_create", "frl_vnet_destroy", "frl_vnet_set_params")

    def __init__(self, state_dim: int, action_dim: int, hidden_dim: int = 256, time_embed_dim: int = 128):
        super().__init__()
        self.state_dim, self.action_dim = state_dim, action_dim
        self.hidden_dim = hidden_dim
        self.time_dim = time_embed_dim
        self.time_embed = nn.Sequential(nn.Linear(1, time_embed_dim), nn.ReLU(),
                                        nn.Linear(time_embed_dim, hidden_dim), nn.ReLU())
        self.state_embed = nn.Sequential(nn.Linear(state_dim, hidden_dim), nn.ReLU())
        self.action_embed = nn.Sequential(nn.Linear(action_dim, hidden_dim), nn.ReLU())
        self.velocity_head = nn.Sequential(nn.Linear(hidden_dim * 3, hidden_dim), nn.ReLU(),
                                           nn.Linear(hidden_dim, action_dim))
        for m in self.modules():          # model.py:419 -> no_bias_weight_init (model.py:14-19)
            if isinstance(m, nn.Linear):
                nn.init.orthogonal_(m.weight.data)
                m.bias.data.fill_(0.0)
        self._finish_init(state_dim, action_dim, hidden_dim)

    def _native_config(self, device_index: int):
        return _native.VNetConfig(self.state_dim, self.action_dim, self.hidden_dim, self.time_dim, device_index)

    def _inputs(self, s, *xs):
        dev = self._flat.device
        s = as_f32_cuda(s, dev, "state")
        B = s.shape[0]
        if s.shape != (B, self.state_dim):
            raise ValueError(f"expected state [B,{self.state_dim}], got {tuple(s.shape)}")
        out = []
        for name, x in xs:
            x = as_f32_cuda(x, dev, name)
            if x.shape != (B, self.action_dim):
                raise ValueError(f"expected {name} [B,{self.action_dim}], got {tuple(x.shape)}")
            out.append(x)
        return (s, *out)

    def _times(self, t, B):
        t = as_f32_cuda(t, self._flat.device, "t").reshape(-1)
        if t.numel() != B:
            raise ValueError(f"expected t [B] or [B,1] with B={B}, got {tuple(t.shape)}")
        return t

    def forward(self, x_t, s, t):
        h = self.native()
        s, x_t = self._inputs(s, ("x_t", x_t))
        B = s.shape[0]
        t = self._times(t, B)
        v = torch.empty(B, self.action_dim, device=s.device)
        _native.check(_native.lib().frl_vnet_forward(h, _ptr(x_t), _ptr(s), _ptr(t), _ptr(v), B, _current_stream(s.device)),
                      "frl_vnet_forward")
        return v

    def sample_actions(self, state, n_steps: int = 32, x0=None):
        """``FlowPolicy.get_action`` (flow_policy.py:49-63): x ~ N(0, I) (drawn here unless ``x0`` is given), then
        ``n_steps`` of ``x = x + v(x, state, i*dt) * dt``."""
        h = self.native()
        dev = self._flat.device
        if x0 is None:
            x0 = torch.randn(state.shape[0], self.action_dim, device=dev)   # flow_policy.py:53
        state, x0 = self._inputs(state, ("x0", x0))
        B = state.shape[0]
        out = torch.empty(B, self.action_dim, device=dev)
        _native.check(_native.lib().frl_vnet_sample(h, _ptr(state), _ptr(x0), int(n_steps), _ptr(out), B,
                                                    _current_stream(dev)), "frl_vnet_sample")
        return out

    def bcf_loss(self, states, actions, x0=None, t=None, bc_coef: float = 1.0, *, grad=True, accumulate=False,
                 mean_divisor: int = 0, grad_scale: float = 1.0):
        """Loss of ``_bc_step`` (bcf.py:108-128) and, with ``grad=True``, its parameter gradient into ``flat_grad()``.
        Returns a device tensor ``[loss, flow_loss, action_loss]``."""
        h = self.native()
        dev = self._flat.device
        if x0 is None:
            x0 = torch.randn_like(actions)                                  # bcf.py:109
        if t is None:
            t = torch.rand(actions.shape[0], 1, device=dev)                 # bcf.py:111
        states, actions, x0 = self._inputs(states, ("actions", actions), ("x0", x0))
        B = states.shape[0]
        t = self._times(t, B)
        losses = torch.empty(3, device=dev)
        g = self.flat_grad() if grad else None
        _native.check(_native.lib().frl_bcf_loss_grad(h, _ptr(states), _ptr(actions), _ptr(x0), _ptr(t), B,
                                                      int(mean_divisor), float(bc_coef), float(grad_scale),
                                                      int(bool(accumulate)), _ptr(losses), _ptr(g),
                                                      _current_stream(dev)), "frl_bcf_loss_grad")
        return losses


class FlowPolicy(nn.Module):
    """The acting side of ``rlf/policies/flow_policy.py``: holds the velocity net, ``forward`` evaluates it and
    ``get_action`` integrates it (``args.flow_integration_steps``, default 32)."""

    def __init__(self, velocity_net: VectorNetwork, flow_integration_steps: int = 32):
        super().__init__()
        self.velocity_net = velocity_net
        self.flow_integration_steps = int(flow_integration_steps)

    def forward(self, x_t, state, t):
        return self.velocity_net(x_t, state, t)

    def get_action(self, state, add_state=None, rnn_hxs=None, mask=None, step_info=None, *, x0=None):
        return self.velocity_net.sample_actions(state, self.flow_integration_steps, x0=x0)


def bc_step(policy, opt: FusedAdam, states, actions, *, x0=None, t=None, bc_coef: float = 1.0, max_grad_norm: float = 0.5):
    """One ``BehaviorCloneFlowMatching._bc_step`` update (bcf.py:99-128): loss, backward, ``clip_grad_norm_`` with
    ``--max-grad-norm`` (default 0.5; <= 0 disables, base_net_algo.py:84-89) and ``Adam.step``.  ``opt`` is a
    ``FusedAdam([net], lr=args.lr, eps=args.eps)`` (rl-toolkit's default eps is 1e-12, base_net_algo.py:130-146).
    Returns the device tensor ``[loss, flow_loss, action_loss]`` (the reference logs their ``.item()``)."""
    net = policy.velocity_net if isinstance(policy, FlowPolicy) else policy
    losses = net.bcf_loss(states, actions, x0=x0, t=t, bc_coef=bc_coef)
    opt.step_flat([net.flat_grad()], max_norm=max_grad_norm if max_grad_norm > 0 else float("inf"))
    return losses
