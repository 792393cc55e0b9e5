"""Flow-matching density estimators of FLOWRIL on B200 kernels.

Drop-in for the estimator classes of the reference's ``modril/flowril/pretrain.py`` (level-1 boundary of SURVEY.md
section 8b): ``CoupledResidualFM`` (:137-214) and ``FlowMatching`` (:62-134) with the same constructor arguments,
attributes (``.net``, ``.prior``, ``.eps``), methods (``v_field``, ``fm_loss``, ``log_prob``) and
``state_dict`` keys; module-level ``_rademacher`` (:42-52).

Mechanism: ``log_prob`` is ONE launch of the fused Euler integrator (``frl_logprob`` / ``frl_score``) that carries
the action state, a forward-mode tangent and the divergence accumulator through all ``n_steps`` on chip;
``fm_loss`` is a fused forward+backward (``frl_cfm_loss_grad``) exposed to autograd through a custom Function, so
``loss.backward()``, ``clip_grad_norm_`` and any torch optimizer keep working; ``FusedAdam`` and
``train_step`` are the all-native route (grad-clip + Adam in one kernel on the flat parameter buffer).

Randomness is drawn on the host side exactly where the reference draws it (same torch calls in the same order),
and every method also accepts the draws explicitly (``probe=``, ``t=``, ``noise=``) -- the kernels themselves are
deterministic functions of their inputs.
"""
from __future__ import annotations

import ctypes as C
import functools
import math

import torch
import torch.nn as nn

from .. import _native
from .networks import ConditionalVNet, SharedVNet, _current_stream, _ptr, as_f32_cuda

CMAP = 'viridis'
PLOT_KW = dict(density=1.2, linewidth=1, arrowstyle='-|>')
ALPHA_BG = 0.85

# "fp16x3": fp32-grade products on the tensor cores (3-term fp16 operand split, csrc/logprob_tc3x.cu) -- the reference
# computes in fp32 (pretrain.py:192-214), and this is the mode that meets the 1e-4 parity bar at tensor-core speed.
_DEFAULT_PRECISION = "fp16x3"


def set_default_precision(name: str):
    """Arithmetic of log_prob / score calls that do not pass ``precision=``: 'fp16x3' (default: tcgen05 tensor cores, 3-term
    fp16 split = fp32-grade, 1e-4 parity), 'fp32' (FFMA kernels, 1e-4 parity), 'fp16' or 'bf16' (single-pass tensor-core
    modes: 2e-3 on default-init-scale / trained nets, looser on rough nets -- README 'Precision modes')."""
    global _DEFAULT_PRECISION
    if name not in _native.PRECISIONS:
        raise ValueError(f"unknown precision {name!r}; choose from {sorted(_native.PRECISIONS)}")
    _DEFAULT_PRECISION = name


def _rademacher(shape_or_tensor, device=None):
    """+-1 probe matching a shape or tensor.  Like the reference (pretrain.py:42-52) it draws from a FRESH,
    unseeded ``torch.Generator`` on every call, so the probe is a deterministic function of (shape, device)."""
    if isinstance(shape_or_tensor, torch.Tensor):
        device = shape_or_tensor.device if device is None else device
        shape = tuple(shape_or_tensor.shape)
    else:
        shape = shape_or_tensor
        assert device is not None, "device must be specified when passing shape"
    rng = torch.Generator(device=device)
    return torch.randint(0, 2, shape, device=device, generator=rng).float().mul_(2).sub_(1)


@functools.lru_cache(maxsize=128)
def _t_mid(n_steps: int, device_type: str = "cuda"):
    """Midpoints of torch.linspace(0, 1, n+1) in float32, formed with torch on the device type the model lives on --
    exactly what the reference does (pretrain.py:200, :206; the CPU and CUDA linspace kernels can differ in the last
    bit).  Cached per (n_steps, device type): one tiny launch + copy the first time only."""
    g = torch.linspace(0.0, 1.0, n_steps + 1, device=device_type)
    tm = 0.5 * (g[:-1] + g[1:])
    return _native.float_array(tm.cpu().tolist())


def _precision_code(precision):
    name = _DEFAULT_PRECISION if precision is None else precision
    if name not in _native.PRECISIONS:
        raise ValueError(f"unknown precision {name!r}")
    return _native.PRECISIONS[name]


def _probe_args(probe, exact, B, a_dim, n_steps, device):
    if exact:
        return None, _native.PROBE_EXACT
    probe = as_f32_cuda(probe, device, "probe")
    if probe.shape == (B, a_dim):
        return probe, _native.PROBE_SHARED
    if probe.shape == (n_steps, B, a_dim):
        return probe, _native.PROBE_PER_STEP
    raise ValueError(f"probe must be [B,{a_dim}] or [{n_steps},B,{a_dim}], got {tuple(probe.shape)}")


def _log_prob_native(net, sign, s, a0, probe, exact, n_steps, precision):
    dev = net.flat_params().device
    h = net.native()
    s = as_f32_cuda(s, dev, "s")
    a0 = as_f32_cuda(a0, dev, "a0")
    B = a0.shape[0]
    if a0.shape != (B, net.a_dim) or s.shape != (B, net.s_dim):
        raise ValueError(f"expected s [B,{net.s_dim}], a0 [B,{net.a_dim}]; got {tuple(s.shape)}, {tuple(a0.shape)}")
    if not (1 <= n_steps <= 256):
        raise ValueError("n_steps must be in 1..256")
    probe, mode = _probe_args(probe, exact, B, net.a_dim, n_steps, dev)
    out = torch.empty(B, device=dev)
    _native.check(_native.lib().frl_logprob(h, float(sign), _ptr(s), _ptr(a0), _ptr(probe), mode, _t_mid(n_steps),
                                            n_steps, _ptr(out), B, _precision_code(precision),
                                            _current_stream(dev)), "frl_logprob")
    return out


def score_pair(net_e, net_a, s, a, probe=None, exact=False, n_steps=32, reward_type="raw", precision=None):
    """Both roles in one launch: returns (logp_E, logp_A, reward).  For scrf pass the same net twice
    (reference flowril.py:183-191); for 2fs the expert and agent nets (flowril.py:187-189)."""
    dev = net_e.flat_params().device
    he, ha = net_e.native(), net_a.native()
    s = as_f32_cuda(s, dev, "s")
    a = as_f32_cuda(a, dev, "a")
    B = a.shape[0]
    if a.shape != (B, net_e.a_dim) or s.shape != (B, net_e.s_dim):
        raise ValueError(f"expected s [B,{net_e.s_dim}], a [B,{net_e.a_dim}]; got {tuple(s.shape)}, {tuple(a.shape)}")
    if reward_type not in _native.REWARD_TYPES:
        raise ValueError(f"Unrecognized reward type {reward_type}")  # reference flowril.py:207
    probe, mode = _probe_args(probe, exact, B, net_e.a_dim, n_steps, dev)
    le, la, rw = (torch.empty(B, device=dev) for _ in range(3))
    _native.check(_native.lib().frl_score(he, ha, _ptr(s), _ptr(a), _ptr(probe), mode, _t_mid(n_steps), n_steps,
                                          _native.REWARD_TYPES[reward_type], _ptr(le), _ptr(la), _ptr(rw), B,
                                          _precision_code(precision), _current_stream(dev)), "frl_score")
    return le, la, rw


def score_pair_host(net_e, net_a, s, a, probe=None, exact=False, n_steps=32, reward_type="raw", precision=None,
                    out=None):
    """End-to-end scoring of HOST tensors (``frl_score_host``): the library stages the batch through the device in
    chunks (H2D, fused integrator, D2H overlapped on its own streams) and returns CPU tensors
    (logp_E, logp_A, reward).  Pinned inputs/outputs (``tensor.pin_memory()``) make the copies asynchronous.
    ``out``: optional tuple of three preallocated CPU float32 tensors [B]."""
    dev = net_e.flat_params().device
    he, ha = net_e.native(), net_a.native()   # (the library synchronises the device itself before it stages the batch)

    def host(x, what):
        if x.device.type != "cpu":
            raise RuntimeError(f"{what} must be a CPU tensor for the host entry point")
        return x.detach().float().contiguous()

    s, a = host(s, "s"), host(a, "a")
    B = a.shape[0]
    if a.shape != (B, net_e.a_dim) or s.shape != (B, net_e.s_dim):
        raise ValueError(f"expected s [B,{net_e.s_dim}], a [B,{net_e.a_dim}]; got {tuple(s.shape)}, {tuple(a.shape)}")
    if reward_type not in _native.REWARD_TYPES:
        raise ValueError(f"Unrecognized reward type {reward_type}")
    if exact:
        probe, mode = None, _native.PROBE_EXACT
    else:
        probe = host(probe, "probe")
        if probe.shape == (B, net_e.a_dim):
            mode = _native.PROBE_SHARED
        elif probe.shape == (n_steps, B, net_e.a_dim):
            mode = _native.PROBE_PER_STEP
        else:
            raise ValueError(f"bad probe shape {tuple(probe.shape)}")
    le, la, rw = out if out is not None else (torch.empty(B) for _ in range(3))
    _native.check(_native.lib().frl_score_host(he, ha, _ptr(s), _ptr(a), _ptr(probe), mode, _t_mid(n_steps), n_steps,
                                               _native.REWARD_TYPES[reward_type], _ptr(le), _ptr(la), _ptr(rw), B,
                                               _precision_code(precision)), "frl_score_host")
    return le, la, rw


class _FMLoss(torch.autograd.Function):
    """loss = frl_cfm_loss_grad(...): the fused kernel produces the loss AND dL/dtheta in forward; backward hands
    the stored gradient (scaled by grad_output) to autograd so that ``loss.backward()`` fills ``p.grad``."""

    @staticmethod
    def forward(ctx, net, sign, need_grad, prec, s, a0, t, noise, *params):
        dev = net.flat_params().device
        h = net.native()
        B = a0.shape[0]
        loss = torch.empty(1, device=dev)
        gbuf = net.native_grad("fmloss") if need_grad else None
        _native.check(_native.lib().frl_cfm_loss_grad_p(h, float(sign), _ptr(s), _ptr(a0), _ptr(t), _ptr(noise), B, 0,
                                                        1.0, 0, _ptr(loss), _ptr(gbuf), prec, _current_stream(dev)),
                      "frl_cfm_loss_grad_p")
        ctx.net = net
        ctx.flat_grad = net.from_native_grad(gbuf).clone() if need_grad else None   # (the buffer is reused by the next call)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        outs = [None] * 8
        flat = ctx.flat_grad
        for p, off, n in ctx.net.param_slices():
            if flat is None or not p.requires_grad:
                outs.append(None)
            else:
                outs.append(flat[off:off + n].view(p.shape) * grad_out)
        return tuple(outs)


_TRAIN_PRECISION = "fp32"


def set_train_precision(name: str):
    """'fp32' (FFMA kernels, the 1e-4 parity mode) or 'fp16' (tcgen05 tensor cores, fp16 operands / fp32 accumulate)
    for ``fm_loss`` / ``train_step`` calls that do not pass ``precision=`` explicitly."""
    global _TRAIN_PRECISION
    if name not in ("fp32", "fp16"):
        raise ValueError("training precision must be 'fp32' or 'fp16'")
    _TRAIN_PRECISION = name


def _train_precision_code(precision):
    name = _TRAIN_PRECISION if precision is None else precision
    if name not in ("fp32", "fp16"):
        raise ValueError("training precision must be 'fp32' or 'fp16'")
    return _native.PRECISIONS[name]


def _fm_loss_native(net, sign, s, a0, t, noise, precision=None):
    dev = net.flat_params().device
    net.native()
    s = as_f32_cuda(s, dev, "s")
    a0 = as_f32_cuda(a0, dev, "a0")
    t = as_f32_cuda(t, dev, "t").reshape(-1)
    noise = as_f32_cuda(noise, dev, "noise")
    B = a0.shape[0]
    if a0.shape != (B, net.a_dim) or s.shape != (B, net.s_dim) or t.numel() != B or noise.shape != a0.shape:
        raise ValueError("fm_loss: inconsistent batch shapes")
    params = [p for p, _, _ in net.param_slices()]
    # (grad mode is always off inside Function.forward, so decide here whether dL/dtheta is wanted)
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    return _FMLoss.apply(net, sign, need_grad, _train_precision_code(precision), s, a0, t, noise, *params)


def reg_losses_native(net, s, a, t, probe, anti=True, stable=True, need_grad=True, mean_divisor=0):
    """Anti-divergence / Jacobian-stability regularisers of the scrf model on one (mixed) batch through
    ``frl_reg_loss_grad`` (reference flowril.py:255-264).  Returns a dict with device scalars ``loss_anti`` /
    ``loss_stable`` and (need_grad) the unscaled flat gradients ``grad_anti`` / ``grad_stable`` (state_dict order;
    buffers owned by the net and overwritten by the next call)."""
    dev = net.flat_params().device
    h = net.native()
    s = as_f32_cuda(s, dev, "s")
    a = as_f32_cuda(a, dev, "a")
    t = as_f32_cuda(t, dev, "t").reshape(-1)
    probe = as_f32_cuda(probe, dev, "probe")
    N = a.shape[0]
    if a.shape != (N, net.a_dim) or s.shape != (N, net.s_dim) or t.numel() != N or probe.shape != a.shape:
        raise ValueError("regularisers: inconsistent batch shapes")
    which = (1 if anti else 0) | (2 if stable else 0)
    if which == 0:
        return {}
    buf = net.native_grad("reg", rows=2)
    losses = torch.zeros(2, device=dev)
    la, ls = losses[0:1], losses[1:2]
    ga = buf[0] if (need_grad and anti) else None
    gs = buf[1] if (need_grad and stable) else None
    _native.check(_native.lib().frl_reg_loss_grad(h, _ptr(s), _ptr(a), _ptr(t), _ptr(probe), N, int(mean_divisor),
                                                  which, _ptr(la if anti else None), _ptr(ls if stable else None),
                                                  _ptr(ga), _ptr(gs), _current_stream(dev)), "frl_reg_loss_grad")
    if need_grad and net.native_hidden() != net.hidden:   # padded width: back to the true layout
        true = getattr(net, "_reg_true", None)
        if true is None or true.device != dev:
            true = torch.zeros(2, net.flat_params().numel(), device=dev)
            net._reg_true = true
        net.from_native_grad(buf, out=true)
        ga = true[0] if anti else None
        gs = true[1] if stable else None
    out = {}
    if anti:
        out["loss_anti"], out["grad_anti"] = la.reshape(()), ga
    if stable:
        out["loss_stable"], out["grad_stable"] = ls.reshape(()), gs
    return out


def tensor_grad_norm_sum(net, flat_grad):
    """sum over the parameter tensors of ||g_tensor||_2 -- the quantity GradNorm2D.update forms with
    ``sum(g.norm() for g in grads)`` (reference networks.py:108-117).  Returns a device scalar."""
    dev = flat_grad.device
    offs = [0]
    for _, off, n in net.param_slices():
        offs.append(off + n)
    arr = (C.c_longlong * len(offs))(*offs)
    out = torch.empty(len(offs) - 1, device=dev)
    _native.check(_native.lib().frl_tensor_norms(_ptr(flat_grad), arr, len(offs) - 1, _ptr(out), dev.index or 0,
                                                 _current_stream(dev)), "frl_tensor_norms")
    return out.sum()


def _axpy(y, x, alpha=1.0, num=None, den=None, eps=0.0, gate=None, coef_out=None):
    dev = y.device
    _native.check(_native.lib().frl_grad_axpy(_ptr(y), _ptr(x), y.numel(), float(alpha), _ptr(num), _ptr(den),
                                              float(eps), _ptr(gate), _ptr(coef_out), dev.index or 0,
                                              _current_stream(dev)), "frl_grad_axpy")


class FlowMatching(nn.Module):
    """One conditional flow net (the '2fs' option uses two of these) -- reference pretrain.py:62-134."""

    def __init__(self, state_dim: int, action_dim: int, hidden_dim: int, depth: int = 4, eps: float = 1e-3):
        super().__init__()
        self.state_dim, self.action_dim = state_dim, action_dim
        self.net = ConditionalVNet(self.state_dim, self.action_dim, hidden_dim, depth)
        self.dim = self.state_dim + self.action_dim
        self.prior = torch.distributions.Normal(0, 1)
        self.eps = eps
        self.T = 1.0

    def _v_star(self, a_t, a0, t, noise):
        return noise - a0

    def fm_loss(self, s, a0, dequant_std=0.02, *, dequant=None, t=None, noise=None, precision=None):
        """CFM loss (reference pretrain.py:94-109).  Draw order when not supplied: dequantisation noise
        (randn_like), t (rand), prior sample -- the reference's order."""
        B = s.size(0)
        dev = a0.device
        if dequant is None:
            dequant = torch.randn_like(a0)
        a0 = a0 + dequant * dequant_std
        if t is None:
            t = torch.rand(B, 1, device=dev) * (1 - 2 * self.eps) + self.eps
        if noise is None:
            noise = self.prior.sample((B, self.action_dim)).to(a0)
        return _fm_loss_native(self.net, 1.0, s, a0, t, noise, precision)

    def log_prob(self, x: torch.Tensor, n_steps: int = 32, *, probe=None, exact=False, precision=None):
        """log p(a|s) for x = cat[s, a] (reference pretrain.py:112-134).  Without ``probe`` a fresh Rademacher
        probe is drawn from the global RNG at every Euler step like ``_hutch_div`` does (pretrain.py:90)."""
        s, a0 = x.split([self.state_dim, x.size(-1) - self.state_dim], dim=-1)
        if probe is None and not exact:
            probe = torch.stack([torch.randint_like(a0, 0, 2).float().mul_(2).sub_(1) for _ in range(n_steps)])
        return _log_prob_native(self.net, 1.0, s, a0, probe, exact, n_steps, precision)


class CoupledResidualFM(nn.Module):
    """v_E = v_c + r, v_pi = v_c - r with r = tanh(head_r)  -- reference pretrain.py:137-214."""

    def __init__(self, state_dim: int, action_dim: int, hidden_dim: int, depth: int = 4, eps: float = 1e-3):
        super().__init__()
        self.s_dim, self.a_dim, self.eps = state_dim, action_dim, eps
        self.net = SharedVNet(state_dim, action_dim, hidden=hidden_dim, depth=depth)
        self.prior = torch.distributions.Normal(0., 1.)

    def v_field(self, a_t, s, t, role: str):
        v_c, r = self.net(a_t, s, t)
        return v_c + r if role == "expert" else v_c - r

    def _v_star(self, a_t, a0, t, noise):
        return noise - a0

    def fm_loss(self, s, a0, role, *, t=None, noise=None, precision=None):
        """CFM loss for one role (reference pretrain.py:179-190); t then noise are drawn like the reference when
        not supplied."""
        B = a0.size(0)
        dev = a0.device
        if t is None:
            t = torch.rand(B, 1, device=dev) * (1 - 2 * self.eps) + self.eps
        if noise is None:
            noise = self.prior.sample((B, self.a_dim)).to(dev)
        return _fm_loss_native(self.net, 1.0 if role == "expert" else -1.0, s, a0, t, noise, precision)

    def log_prob(self, s, a0, role: str, n_steps: int = 32, *, probe=None, exact=False, precision=None):
        """log p_role(a|s) (reference pretrain.py:192-214).  Default probe = ``_rademacher((B, a_dim))``, the same
        matrix at every step and for both roles, as in the reference."""
        if probe is None and not exact:
            probe = _rademacher((a0.size(0), self.a_dim), device=a0.device)
        return _log_prob_native(self.net, 1.0 if role == "expert" else -1.0, s, a0, probe, exact, n_steps, precision)

    def reg_losses(self, s, a, *, t=None, probe=None, anti=True, stable=True, need_grad=False):
        """(loss_anti, loss_stable) of reference flowril.py:259-263 on the batch (s, a): t ~ U[0,1) and the
        shape-deterministic ``_rademacher`` probe are drawn like the reference when not supplied."""
        if t is None:
            t = torch.rand(a.size(0), 1, device=a.device)
        if probe is None:
            probe = _rademacher((a.size(0), self.a_dim), device=a.device)
        return reg_losses_native(self.net, s, a, t, probe, anti, stable, need_grad)

    def score(self, s, a, n_steps: int = 32, reward_type: str = "raw", *, probe=None, exact=False, precision=None):
        """(logp_E, logp_A, reward) in one launch -- what flowril.py:183-207 computes with two log_prob calls."""
        if probe is None and not exact:
            probe = _rademacher((a.size(0), self.a_dim), device=a.device)
        return score_pair(self.net, self.net, s, a, probe, exact, n_steps, reward_type, precision)


# ------------------------------------------------------------------------------------------------------------
# optimiser: torch.optim.Adam semantics on the flat buffers, one fused clip+Adam kernel
# ------------------------------------------------------------------------------------------------------------
class FusedAdam(torch.optim.Optimizer):
    """Adam (torch defaults: betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad) whose ``step`` is the
    fused ``frl_clip_adam`` kernel.  ``state_dict()`` / ``load_state_dict()`` use ``torch.optim.Adam``'s layout
    (per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``), so reference checkpoints (``flow_<option>_opt``,
    reference flowril.py:581-597) round-trip.

    ``nets`` are the ``_FlatVNet`` modules whose parameters are optimised; parameters with
    ``requires_grad == False`` at construction are frozen (excluded from the norm and the update, like params with
    ``grad is None`` in ``clip_grad_norm_`` / Adam).
    """

    def __init__(self, nets, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, capturable=False):
        """capturable=True keeps the step count on the device (``frl_clip_adam_dev``), so a whole update
        (``train_step``) can be captured into a CUDA graph (``torch.cuda.graph``) and replayed."""
        self.nets = list(nets)
        params = [p for net in self.nets for p, _, _ in net.param_slices() if p.requires_grad]
        if not params:
            raise ValueError("optimizer got an empty parameter list")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False,
                                      foreach=None, capturable=bool(capturable), differentiable=False, fused=None))
        self._step = 0
        self._step_dev = None
        if capturable:
            self._step_dev = torch.zeros(1, dtype=torch.int64, device=self.nets[0].flat_params().device)
        self._m = [torch.zeros_like(net.flat_params()) for net in self.nets]
        self._v = [torch.zeros_like(net.flat_params()) for net in self.nets]
        self._norm = None
        self._seg_cache = None
        self._bind_state()

    def step_count(self) -> int:
        if self._step_dev is not None:
            self._step = int(self._step_dev.item())
        return self._step

    def _bind_state(self):
        for net, m, v in zip(self.nets, self._m, self._v):
            for p, off, n in net.param_slices():
                if p.requires_grad:
                    self.state[p] = {"step": torch.tensor(float(self._step)),
                                     "exp_avg": m[off:off + n].view(p.shape),
                                     "exp_avg_sq": v[off:off + n].view(p.shape)}

    def state_dict(self):
        step = float(self.step_count())
        for st in self.state.values():   # the per-parameter step tensors are refreshed lazily, not every step
            st["step"].fill_(step)
        return super().state_dict()

    def _segments(self, grads):
        """Contiguous runs of trainable parameters per net -> frl_adam_segment array (cached while the buffers and
        the trainable pattern stay the same)."""
        key = tuple((net.flat_params().data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(),
                     tuple(p.requires_grad for p, _, _ in net.param_slices()))
                    for net, m, v, g in zip(self.nets, self._m, self._v, grads))
        if self._seg_cache is not None and self._seg_cache[0] == key:
            return self._seg_cache[1], self._seg_cache[2]
        segs = []
        for net, m, v, g in zip(self.nets, self._m, self._v, grads):
            flat = net.flat_params()
            run = None
            for p, off, n in net.param_slices() + [(None, None, None)]:
                if p is not None and p.requires_grad:
                    run = [run[0], off + n] if run else [off, off + n]
                elif run:
                    segs.append((flat, g, m, v, run[0], run[1] - run[0]))
                    run = None
        if len(segs) > 4:
            raise NotImplementedError("more than 4 disjoint trainable parameter ranges")
        arr = (_native.AdamSegment * len(segs))()
        for i, (flat, g, m, v, off, n) in enumerate(segs):
            arr[i].params = flat.data_ptr() + 4 * off
            arr[i].grad = g.data_ptr() + 4 * off
            arr[i].exp_avg = m.data_ptr() + 4 * off
            arr[i].exp_avg_sq = v.data_ptr() + 4 * off
            arr[i].n = n
        self._seg_cache = (key, arr, len(segs))
        return arr, len(segs)

    @torch.no_grad()
    def step_flat(self, grads, max_norm: float = float("inf")):
        """Clip (global norm over all trainable segments) + Adam on flat gradient buffers, one kernel pair."""
        dev = self.nets[0].flat_params().device
        group = self.param_groups[0]
        arr, n = self._segments(grads)
        if self._norm is None or self._norm.device != dev:
            self._norm = torch.zeros(1, device=dev)
        b1, b2 = group["betas"]
        mx = float(max_norm) if math.isfinite(max_norm) else 3.0e38
        L = _native.lib()
        if self._step_dev is not None:
            _native.check(L.frl_clip_adam_dev(arr, n, _ptr(self._step_dev), float(group["lr"]), float(b1), float(b2),
                                              float(group["eps"]), mx, _ptr(self._norm), dev.index or 0,
                                              _current_stream(dev)), "frl_clip_adam_dev")
        else:
            self._step += 1
            _native.check(L.frl_clip_adam(arr, n, self._step, float(group["lr"]), float(b1), float(b2),
                                          float(group["eps"]), mx, _ptr(self._norm), dev.index or 0,
                                          _current_stream(dev)), "frl_clip_adam")
        for net in self.nets:
            net.mark_dirty()
        return self._norm

    @torch.no_grad()
    def step(self, closure=None):
        """torch-style step: consumes ``p.grad`` (already clipped by the caller, if at all)."""
        loss = closure() if closure is not None else None
        grads = []
        for net in self.nets:
            g = net.flat_grad()
            g.zero_()
            for p, off, n in net.param_slices():
                if p.requires_grad and p.grad is not None:
                    g[off:off + n].copy_(p.grad.reshape(-1))
            grads.append(g)
        self.step_flat(grads)
        return loss

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # torch replaced the state tensors by copies: move them back into the flat buffers
        steps = []
        for net, m, v in zip(self.nets, self._m, self._v):
            for p, off, n in net.param_slices():
                st = self.state.get(p)
                if st:
                    m[off:off + n].copy_(st["exp_avg"].reshape(-1))
                    v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                    steps.append(int(float(st["step"])))
        self._step = max(steps) if steps else 0
        if self._step_dev is not None:
            self._step_dev.fill_(self._step)
        self._seg_cache = None
        self._bind_state()


def train_step(terms, opt: FusedAdam, max_norm: float = 1.0, world=None, reg=None, precision=None, phase=None,
               losses=None):
    """One reward-model update on the all-native route (reference flowril.py:311-330):
    zero grads; for each term accumulate rate * d fm_loss / d theta; [regularisers]; [all-reduce];
    clip_grad_norm_(max_norm); Adam.

    terms: list of dicts {net, sign, s, a0, t, noise, rate}; several terms may share a net (scrf: expert + agent).
    world: optional (process_group, world_size[, global_rows]) for data-parallel training: each rank passes its shard,
           the mean divisor becomes the global batch (``global_rows`` per term when the shards are ragged, else
           local rows * world_size) and the flat gradients are all-reduced (sum) before the clip.  ``process_group`` may
           be a ``modril_b200.dist.PeerAllReduce``: the all-reduce then runs as two kernels on the current stream over
           NVLink peer memory and the whole update (phase=None) can be captured as ONE CUDA graph.
    reg:   optional dict for the scrf regularisers (reference flowril.py:255-290):
             net, s, a, t, probe      the mixed batch, its times and Rademacher probe
             anti / stable            how each term is weighted: None (off), a float (fixed weight,
                                      flowril.py:266-270), ("warmup", scale, eps) for w = L_E/(L_reg+eps)*scale with
                                      L_E = the first term's loss (flowril.py:272-284), or a device scalar tensor
                                      (GradNorm2D's beta / gamma, networks.py:104-106)
             gate_anti / gate_stable  optional 2-float device tensors holding the previous weight: a weight <= 0
                                      switches the term off (``use_anti = args.loss_anti > 0``) and the new weight is
                                      written back into them
           On return reg["out"] holds loss_anti / loss_stable (device scalars) and grad_anti / grad_stable.
    precision: 'fp32' (FFMA) or 'fp16' (tcgen05 tensor cores) for the CFM terms; default: ``set_train_precision``.
    phase: None runs the whole update.  'grad' (losses + gradients), 'reduce' (the all-reduce) and 'apply' (regulariser
           weights, clip, Adam) run one part each, so that a data-parallel loop can capture 'grad' and 'apply' into CUDA
           graphs and keep the NCCL call between them outside of any graph; pass the 'grad' phase's return value as
           ``losses=`` to the later phases.
    Returns the per-term loss tensors (device scalars; under DP the local shard's contribution summed over ranks).
    """
    if phase not in (None, "grad", "reduce", "apply"):
        raise ValueError(f"unknown phase {phase!r}")
    ws = 1 if world is None else world[1]
    if phase in (None, "grad"):
        losses = _train_step_grad(terms, opt, world, reg, precision)
        if phase == "grad":
            return [l.reshape(()) for l in losses]
    losses = [] if losses is None else list(losses)
    grads = [net.flat_grad() for net in opt.nets]
    rout = reg.get("out") if reg is not None else None
    if phase in (None, "reduce") and world is not None and ws > 1:
        from ..dist import PeerAllReduce, allreduce_flat, allreduce_peer

        extra_g = [] if rout is None else [rout[k] for k in ("grad_anti", "grad_stable") if k in rout]
        extra_l = [] if rout is None else [rout[k] for k in ("loss_anti", "loss_stable") if k in rout]
        if isinstance(world[0], PeerAllReduce):   # two kernels on this stream (graph-capturable), no NCCL launch
            allreduce_peer(world[0], grads + extra_g, losses + extra_l)
        else:
            allreduce_flat(grads + extra_g, losses + extra_l, world[0])
    if phase == "reduce":
        return [l.reshape(()) for l in losses]
    if rout is not None:
        target = reg["net"].flat_grad()
        for key in ("anti", "stable"):
            spec = reg.get(key)
            if spec is None:
                continue
            g_reg, l_reg, gate = rout["grad_" + key], rout["loss_" + key], reg.get("gate_" + key)
            if isinstance(spec, torch.Tensor):
                _axpy(target, g_reg, 1.0, num=spec.reshape(-1), gate=gate, coef_out=gate)
            elif isinstance(spec, tuple) and spec[0] == "warmup":
                _axpy(target, g_reg, float(spec[1]), num=losses[0], den=l_reg.reshape(-1), eps=float(spec[2]),
                      gate=gate, coef_out=gate)
            else:
                _axpy(target, g_reg, float(spec), gate=gate, coef_out=gate)
    opt.step_flat(grads, max_norm)
    return [l.reshape(()) for l in losses]


class TrainStepGraph:
    """One whole update -- ``train_step(terms, opt, ...)``: gradient kernels, [device-side all-reduce], clip, Adam, weight
    re-pack -- captured ONCE as a CUDA graph and replayed per call.  The tensors inside ``terms`` / ``reg`` are the
    graph's static inputs: copy each new batch into them (``copy_``) before calling.  Replays advance the weights through
    raw pointers, which torch's version counters do not see, so every call marks the nets dirty: scoring launched afterwards
    re-packs its own weight layout instead of silently using the one from before the update.

    ``opt`` must be ``FusedAdam(..., capturable=True)`` (step count on the device).  Data parallel: pass
    ``world=(PeerAllReduce, world_size[, global_rows])``; an NCCL process group cannot be captured (use ``train_step``'s
    phases around the collective instead).  Construction runs one eager update to size the workspaces (allocation is
    illegal during capture) and then restores weights and optimizer state, so it has no effect on training."""

    def __init__(self, terms, opt: FusedAdam, max_norm: float = 1.0, world=None, reg=None, precision=None):
        if opt._step_dev is None:
            raise ValueError("TrainStepGraph needs FusedAdam(..., capturable=True)")
        if world is not None and world[1] > 1:
            from ..dist import PeerAllReduce

            if not isinstance(world[0], PeerAllReduce):
                raise ValueError("TrainStepGraph: a captured data-parallel update needs world[0] = PeerAllReduce")
        self.opt, self.terms = opt, terms
        dev = opt.nets[0].flat_params().device
        saved = [(net.flat_params().clone(), m.clone(), v.clone()) for net, m, v in zip(opt.nets, opt._m, opt._v)]
        step = opt._step_dev.clone()
        run = lambda: train_step(terms, opt, max_norm, world=world, reg=reg, precision=precision)
        run()
        with torch.no_grad():
            for net, m, v, (p0, m0, v0) in zip(opt.nets, opt._m, opt._v, saved):
                net.flat_params().copy_(p0)
                m.copy_(m0)
                v.copy_(v0)
                net.mark_dirty()
            opt._step_dev.copy_(step)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.losses = run()
        with torch.no_grad():   # (capture does not execute, but the host-side dirty flags were touched)
            for net in opt.nets:
                net.mark_dirty()

    def __call__(self):
        self.graph.replay()
        for net in self.opt.nets:
            net.mark_dirty()
        return self.losses


def _train_step_grad(terms, opt, world, reg, precision):
    """Losses and gradients of one update (the part of ``train_step`` before the all-reduce); the regulariser outputs go
    to reg["out"]."""
    L = _native.lib()
    prec = _train_precision_code(precision)
    seen = []
    losses = []
    ws = 1 if world is None else world[1]
    global_rows = world[2] if (world is not None and len(world) > 2) else None
    def inputs(term, dev):
        return (as_f32_cuda(term["s"], dev, "s"), as_f32_cuda(term["a0"], dev, "a0"),
                as_f32_cuda(term["t"], dev, "t").reshape(-1), as_f32_cuda(term["noise"], dev, "noise"))

    i = 0
    while i < len(terms):
        term = terms[i]
        net = term["net"]
        dev = net.flat_params().device
        h = net.native()
        padded = net.native_hidden() != net.hidden
        g = net.native_grad("step") if padded else net.flat_grad()   # (padded width: gathered into flat_grad() below)
        first = all(net is not n for n in seen)
        if first:
            seen.append(net)
        # two consecutive terms on the same net (scrf: expert + agent batch) go through the net in ONE stacked pass on
        # the tensor-core route
        pair = prec != _native.PREC_FP32 and i + 1 < len(terms) and terms[i + 1]["net"] is net
        group = terms[i:i + 2] if pair else [term]
        keep, out = [], []
        arr = (_native.CfmTerm * 2)()
        for k, tm in enumerate(group):
            s, a0, t, noise = inputs(tm, dev)
            B = a0.shape[0]
            loss = torch.empty(1, device=dev)
            keep.append((s, a0, t, noise))
            out.append(loss)
            arr[k].role_sign = float(tm["sign"])
            arr[k].s, arr[k].a0, arr[k].t, arr[k].noise = s.data_ptr(), a0.data_ptr(), t.data_ptr(), noise.data_ptr()
            arr[k].B = B
            arr[k].mean_divisor = int(global_rows) if global_rows else B * ws
            arr[k].grad_scale = float(tm.get("rate", 1.0))
            arr[k].loss = loss.data_ptr()
        if pair:
            _native.check(L.frl_cfm_loss_grad_pair(h, arr, 0 if first else 1, _ptr(g), prec, _current_stream(dev)),
                          "frl_cfm_loss_grad_pair")
        else:
            s, a0, t, noise = keep[0]
            _native.check(L.frl_cfm_loss_grad_p(h, arr[0].role_sign, _ptr(s), _ptr(a0), _ptr(t), _ptr(noise), arr[0].B,
                                                arr[0].mean_divisor, arr[0].grad_scale, 0 if first else 1, _ptr(out[0]),
                                                _ptr(g), prec, _current_stream(dev)), "frl_cfm_loss_grad_p")
        losses.extend(out)
        i += len(group)
    grads = []
    for net in opt.nets:
        g = net.flat_grad()
        if all(net is not n for n in seen):
            g.zero_()
        elif net.native_hidden() != net.hidden:
            net.from_native_grad(net.native_grad("step"), out=g)
        grads.append(g)
    rout = None
    if reg is not None and (reg.get("anti") is not None or reg.get("stable") is not None):
        rnet = reg["net"]
        N = reg["a"].shape[0]
        rout = reg_losses_native(rnet, reg["s"], reg["a"], reg["t"], reg["probe"], anti=reg.get("anti") is not None,
                                 stable=reg.get("stable") is not None, need_grad=True,
                                 mean_divisor=int(reg["global_rows"]) if reg.get("global_rows") else N * ws)
    if reg is not None:
        reg["out"] = rout
    return losses


# ------------------------------------------------------------------------------------------------------------
# offline pre-training (reference pretrain.py:217-526, the ``__main__`` of the module) with a device-resident sampler
# ------------------------------------------------------------------------------------------------------------
def norm_vec(x, mean, std):
    """``clamp((x - mean) / (std + 1e-8), -10, 10)`` (reference pretrain.py:33-39) as one ``frl_norm_vec`` launch."""
    dev = x.device
    x = as_f32_cuda(x, dev, "x")
    mean, std = as_f32_cuda(mean, dev, "mean").reshape(-1), as_f32_cuda(std, dev, "std").reshape(-1)
    if x.dim() != 2 or mean.numel() != x.shape[1] or std.numel() != x.shape[1]:
        raise ValueError("norm_vec expects x [n,d] and mean / std [d]")
    out = torch.empty_like(x)
    _native.check(_native.lib().frl_norm_vec(_ptr(x), _ptr(mean), _ptr(std), x.shape[0], x.shape[1], _ptr(out),
                                             dev.index or 0, _current_stream(dev)), "frl_norm_vec")
    return out


def column_moments(x):
    """``x.mean(0), x.std(0)`` (unbiased) of a CUDA matrix [n,d] in one ``frl_column_moments`` launch."""
    dev = x.device
    x = as_f32_cuda(x, dev, "x")
    mean, std = torch.empty(x.shape[1], device=dev), torch.empty(x.shape[1], device=dev)
    _native.check(_native.lib().frl_column_moments(_ptr(x), x.shape[0], x.shape[1], _ptr(mean), _ptr(std), dev.index or 0,
                                                   _current_stream(dev)), "frl_column_moments")
    return mean, std


class _TorchDraws:
    """The random draws of one pre-training batch, from torch's global RNG in the reference's order (pretrain.py:368-378:
    fm_loss(expert) -> rand, Normal.sample; randn_like(a); fm_loss(agent) -> rand, Normal.sample; rand_like(t_noise))."""

    def __init__(self, device, a_dim, eps):
        self.device, self.a_dim, self.eps = device, a_dim, eps

    def perm(self, n):                       # DataLoader(shuffle=True)
        return torch.randperm(n, device=self.device)

    def cfm(self, B):
        t = torch.rand(B, 1, device=self.device) * (1 - 2 * self.eps) + self.eps
        return t, torch.randn(B, self.a_dim, device=self.device)

    def normal(self, B):                     # randn_like(a): the agent perturbation (scrf) / the dequantisation (2fs)
        return torch.randn(B, self.a_dim, device=self.device)

    def t_mix(self, B):
        return torch.rand(B, 1, device=self.device)


class OfflinePretrainer:
    """The training loop of the reference's offline pre-trainer (pretrain.py:269-440) with the demonstrations resident in
    HBM: optional ``norm_vec`` of observations and actions with their own column moments (:273-285), the odd-row rule
    (:289-290), shuffled 128-row batches (:297; the last one ragged), and per batch

      scrf: ``fm_loss(s, a, expert) + fm_loss(s, a + 0.1 randn, agent)`` (:368-372) plus the anti-divergence /
            Jacobian-stability terms on ``(s, a_noise)`` (:374-381) with fixed weights (``autotune`` false, :383-387) or the
            warm-up weights ``L_E / (L_reg + 1e-8) * 1e-3`` (:391-402; the reference never leaves this branch because
            ``gradnorm(...)`` is only called after it), ``clip_grad_norm_(1)`` and Adam (:417-420);
      2fs:  ``FlowMatching.fm_loss(s, a)`` (:414-415).

    Every batch is one ``train_step`` (fused loss + gradient kernels, clip + Adam kernel); nothing is copied to the host
    inside an epoch.  ``draws`` (tests) replaces the torch RNG: an object with ``perm / cfm / normal / t_mix``."""

    def __init__(self, obs, actions, *, option="scrf", hidden_dim=128, depth=4, eps=1e-2, lr=1e-4, norm=True,
                 enable_loss_anti=True, enable_loss_stable=True, autotune=True, loss_anti=1e-4, loss_stable=1e-6,
                 batch_size=128, device="cuda", precision=None):
        if option not in ("scrf", "2fs"):
            raise NotImplementedError(option)
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("the pre-trainer computes only on CUDA (sm_100a) devices; there is no CPU fallback")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        obs = torch.as_tensor(obs).to(dev, torch.float32).contiguous()
        actions = torch.as_tensor(actions).to(dev, torch.float32).contiguous()
        if obs.dim() != 2 or actions.dim() != 2 or obs.shape[0] != actions.shape[0]:
            raise ValueError("obs [n,s_dim] and actions [n,a_dim] must have the same number of rows")
        self.obs_moments = self.action_moments = None
        if norm:
            self.obs_moments = column_moments(obs)
            obs = norm_vec(obs, *self.obs_moments)
            self.action_moments = column_moments(actions)
            actions = norm_vec(actions, *self.action_moments)
        if obs.shape[0] % 2 == 1:                                    # pretrain.py:289-290
            obs, actions = obs[1:].contiguous(), actions[1:].contiguous()
        self.obs, self.actions = obs, actions
        self.s_dim, self.a_dim = obs.shape[1], actions.shape[1]
        self.option, self.batch_size, self.precision, self.device = option, int(batch_size), precision, dev
        self.autotune = bool(autotune)
        if option == "scrf":
            self.estimator = CoupledResidualFM(self.s_dim, self.a_dim, hidden_dim, depth=depth, eps=eps).to(dev)
        else:
            self.estimator = FlowMatching(self.s_dim, self.a_dim, hidden_dim, depth=depth, eps=eps).to(dev)
        self.estimator.train()
        self.opt = FusedAdam([self.estimator.net], lr=lr)
        w_a = float(loss_anti) if enable_loss_anti else 0.0          # pretrain.py:336-339
        w_s = float(loss_stable) if enable_loss_stable else 0.0
        self._w_init = (w_a, w_s)
        self._w_anti = torch.tensor([w_a, 0.0], device=dev)          # args.loss_anti / args.loss_stable on the device
        self._w_stable = torch.tensor([w_s, 0.0], device=dev)
        self.train_loss_list = []
        self.steps = 0

    # one batch: rows `i` of the resident dataset
    def _step(self, i, draws):
        s, a = self.obs[i], self.actions[i]
        B = a.shape[0]
        net = self.estimator.net
        if self.option == "2fs":
            a0 = a + draws.normal(B) * 0.02                          # FlowMatching.fm_loss dequantisation (:98)
            t, noise = draws.cfm(B)
            losses = train_step([dict(net=net, sign=1.0, s=s, a0=a0, t=t, noise=noise, rate=1.0)], self.opt, max_norm=1.0,
                                precision=self.precision)
            return losses[0], losses[0]
        tE, nE = draws.cfm(B)
        a_noise = a + 0.1 * draws.normal(B)                          # :370-371
        tA, nA = draws.cfm(B)
        terms = [dict(net=net, sign=1.0, s=s, a0=a, t=tE, noise=nE, rate=1.0),
                 dict(net=net, sign=-1.0, s=s, a0=a_noise, t=tA, noise=nA, rate=1.0)]
        reg = None
        if self._w_init[0] > 0 or self._w_init[1] > 0:               # use_anti / use_stable (:374-376)
            spec = ("warmup", 1e-3, 1e-8) if self.autotune else None
            reg = dict(net=net, s=s, a=a_noise, t=draws.t_mix(B), probe=_rademacher((B, self.a_dim), device=self.device),
                       anti=(spec or self._w_init[0]) if self._w_init[0] > 0 else None,
                       stable=(spec or self._w_init[1]) if self._w_init[1] > 0 else None,
                       gate_anti=self._w_anti, gate_stable=self._w_stable)
        lE, lA = train_step(terms, self.opt, max_norm=1.0, reg=reg, precision=self.precision)
        loss = lE + lA
        if reg is not None and reg.get("out"):
            out = reg["out"]
            if "loss_anti" in out:
                loss = loss + self._w_anti[0] * out["loss_anti"].reshape(())
            if "loss_stable" in out:
                loss = loss + self._w_stable[0] * out["loss_stable"].reshape(())
        return loss, lE

    def epoch(self, draws=None):
        """One pass over the demonstrations (pretrain.py:362-440); returns the device tensor of per-batch
        ``[loss, loss_expert]`` rows and appends the epoch's average loss to ``train_loss_list``."""
        draws = draws or _TorchDraws(self.device, self.a_dim, self.estimator.eps)
        n = self.obs.shape[0]
        perm = draws.perm(n)
        rows = []
        for k in range(0, n, self.batch_size):
            loss, l_e = self._step(perm[k:k + self.batch_size], draws)
            rows.append(torch.stack([loss.reshape(()), l_e.reshape(())]))
            self.steps += 1
        out = torch.stack(rows)
        self.train_loss_list.append(float(out[:, 0].mean()))        # the epoch's one device->host copy
        return out

    def fit(self, num_epoch, log_every=0):
        for t in range(1, num_epoch + 1):
            self.epoch()
            if log_every and t % log_every == 0:
                print(f"epoch {t}: loss {self.train_loss_list[-1]:.6f}", flush=True)
        return self.train_loss_list

    def save(self, save_path, env, task):
        """The reference's files (pretrain.py:520-526): ``<save_path>/<env>/<task>/trained_models/<env>_fm.pt`` (the
        estimator's ``state_dict``, loadable through ``--flow-path``) and ``<env>_train_loss.pt``."""
        import os

        d = os.path.join(save_path, env, task, "trained_models")
        os.makedirs(d, exist_ok=True)
        torch.save({k: v.detach().cpu() for k, v in self.estimator.state_dict().items()}, os.path.join(d, f"{env}_fm.pt"))
        torch.save(self.train_loss_list, os.path.join(d, f"{env}_train_loss.pt"))
        return d


def main(argv=None):
    """``python -m modril_b200.flowril.pretrain`` -- the reference's command line (pretrain.py:217-262), same flags."""
    import argparse

    str2bool = lambda v: v.lower() == "true"
    parser = argparse.ArgumentParser()
    parser.add_argument('--traj-load-path', type=str, default='modril/expert_datasets/sine.pt')
    parser.add_argument('--hidden-dim', type=int, default=128)
    parser.add_argument('--depth', type=int, default=4)
    parser.add_argument('--norm', type=str2bool, default=True)
    parser.add_argument('--lr', type=float, default=1e-4)
    parser.add_argument('--eps', type=float, default=1e-2)
    parser.add_argument('--option', type=str, default='scrf', choices=['scrf', '2fs'])
    parser.add_argument('--num-epoch', type=int, default=8000)
    parser.add_argument('--prefix', type=str, default='')
    parser.add_argument('--save-path', type=str, default='data/pre')
    parser.add_argument('--enable-loss-anti', type=str2bool, default=True)
    parser.add_argument('--enable-loss-stable', type=str2bool, default=True)
    parser.add_argument('--autotune', type=str2bool, default=True)
    parser.add_argument('--loss-anti', type=float, default=1e-4)
    parser.add_argument('--loss-stable', type=float, default=1e-6)
    args = parser.parse_args(argv)
    task = args.option                                               # pretrain.py:240-251
    if not args.enable_loss_anti and args.enable_loss_stable:
        task = f"{task}_no_a"
    if not args.enable_loss_stable and args.enable_loss_anti:
        task = f"{task}_no_s"
    if not args.enable_loss_anti and not args.enable_loss_stable:
        task = f"{task}_no_as"
    env = args.traj_load_path.split('/')[-1][:-3]
    if env.lower() == "sine":
        args.norm = False                                            # :258-259
    data = torch.load(args.traj_load_path)
    trainer = OfflinePretrainer(data["obs"], data["actions"], option=args.option, hidden_dim=args.hidden_dim, depth=args.depth,
                                eps=args.eps, lr=args.lr, norm=args.norm, enable_loss_anti=args.enable_loss_anti,
                                enable_loss_stable=args.enable_loss_stable, autotune=args.autotune, loss_anti=args.loss_anti,
                                loss_stable=args.loss_stable)
    print(f"Task = {task}\nHidden dimension = {args.hidden_dim}\nDepth = {args.depth}\nobs {tuple(trainer.obs.shape)} "
          f"actions {tuple(trainer.actions.shape)}")
    trainer.fit(args.num_epoch, log_every=max(1, args.num_epoch // 20))
    print("saved to", trainer.save(args.save_path, env, task))


if __name__ == "__main__":
    main()
