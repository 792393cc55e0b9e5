"""Velocity-field MLPs of the FLOWRIL density model, backed by the sm_100a kernels.

Drop-in for the reference's ``modril/flowril/networks.py``: same class names, constructor arguments, attribute
names (``.shared``, ``.head_c``, ``.head_r``, ``.net``) and ``state_dict`` keys, so
pre-trained ``*_fm.pt`` files load unchanged and ``flowril.py``'s freezing code (``net.shared.parameters()`` ...,
reference flowril.py:52-58) keeps working.

Differences in mechanism (not in behaviour):
  * all parameters of one net are views into ONE flat float32 buffer in ``state_dict`` order; the C library packs
    its kernel-side layouts (transposed / padded / bf16) from that buffer (``frl_net_set_params``) whenever a
    parameter's version counter shows it changed;
  * ``forward`` calls ``frl_vfield``; there is no PyTorch implementation of the arithmetic in this package and no
    CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from .. import _native


def _current_stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def as_f32_cuda(x: torch.Tensor, device, what: str) -> torch.Tensor:
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{what} must be a torch.Tensor")
    if x.device != device:
        raise RuntimeError(f"{what} is on {x.device}, the flow net is on {device}")
    if x.dtype != torch.float32:
        x = x.float()
    return x.detach().contiguous()


class _FlatVNet(nn.Module):
    """Common machinery: flat parameter storage + native handle."""

    n_heads = 0

    def _finish_init(self, s_dim: int, a_dim: int, hidden: int):
        self.s_dim, self.a_dim, self.hidden = int(s_dim), int(a_dim), int(hidden)
        self._flat = None
        self._flat_grad = None
        self._handle = None
        self._packed_key = None
        self._pad = None
        self._flatten()

    # ---- free --hidden-dim (reference flowril.py:562, pretrain CLI): the kernels are built for trunk widths 128 and 256;
    # any other width <= 256 runs as the next of the two with the extra units' weights and biases held at ZERO --
    # relu(0) = 0 feeds nothing forward, so values, tangents and the gradients of the real entries are unchanged.  The
    # nn.Parameters keep the reference's shapes; the library sees a zero-padded image of the flat vector.
    def native_hidden(self) -> int:
        return self.hidden

    def _pad_plan(self):
        """(index of every true flat entry inside the padded flat vector, padded length), or None without padding."""
        Hp = self.native_hidden()
        if Hp == self.hidden:
            return None
        pad = self.__dict__.get("_pad")
        dev = self._flat.device
        if pad is not None and pad[0].device == dev:
            return pad
        H, A, d_in = self.hidden, self.a_dim, self.a_dim + self.s_dim + 1
        ar = lambda n: torch.arange(n, device=dev)
        parts, off = [], 0
        for l in range(self.depth):
            K, Kp = (d_in, d_in) if l == 0 else (H, Hp)
            parts += [(ar(H)[:, None] * Kp + ar(K)[None, :] + off).reshape(-1), ar(H) + off + Hp * Kp]
            off += Hp * Kp + Hp
        for _ in range(self.n_heads):
            parts += [(ar(A)[:, None] * Hp + ar(H)[None, :] + off).reshape(-1), ar(A) + off + A * Hp]
            off += A * Hp + A
        idx = torch.cat(parts)
        assert idx.numel() == self._flat.numel()
        self.__dict__["_pad"] = (idx, off, torch.zeros(off, device=dev), {})
        return self.__dict__["_pad"]

    def native_grad(self, key="grad", rows: int = 0):
        """The buffer to hand to the library for a flat gradient: the caller's own true-size buffer for native widths,
        a persistent zero-padded one otherwise (``rows`` > 0: a [rows, n] block)."""
        pad = self._pad_plan()
        n = self._flat.numel() if pad is None else pad[1]
        cache = self.__dict__.setdefault("_native_grads", {})
        buf = cache.get(key)
        shape = (rows, n) if rows else (n,)
        if buf is None or buf.device != self._flat.device or tuple(buf.shape) != shape:
            buf = torch.zeros(shape, device=self._flat.device)
            cache[key] = buf
        return buf

    def from_native_grad(self, buf, out=None):
        """True-layout view of a gradient the library wrote into a ``native_grad`` buffer (a gather for padded nets)."""
        pad = self._pad_plan()
        if pad is None:
            if out is not None and out.data_ptr() != buf.data_ptr():
                out.copy_(buf)
            return buf if out is None else out
        got = buf.index_select(buf.dim() - 1, pad[0])
        if out is None:
            return got
        out.copy_(got)
        return out

    def __getstate__(self):
        st = self.__dict__.copy()
        st.update(_handle=None, _packed_key=None, _flat=None, _flat_grad=None, _plist=None, _pad=None,
                  _native_grads={})  # native handle and device buffers are per-object
        return st

    def __setstate__(self, st):
        super().__setstate__(st)
        self._flatten()

    # ---- flat storage -------------------------------------------------------------------------------------
    def _ordered_params(self):
        # registration order == state_dict order.  The module tree is fixed after __init__, so the list is cached
        # (walking nn.Module.parameters() costs ~30 us, paid several times per native() call otherwise).
        plist = self.__dict__.get("_plist")
        if plist is None:
            plist = list(self.parameters())
            self.__dict__["_plist"] = plist
        return plist

    def _flatten(self):
        self.__dict__["_plist"] = None
        params = self._ordered_params()
        dev = params[0].device
        if any(p.dtype != torch.float32 for p in params):
            raise TypeError("flow nets are float32 only")
        flat = torch.empty(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in params:
                n = p.numel()
                flat[off:off + n].copy_(p.detach().reshape(-1))
                p.data = flat[off:off + n].view(p.shape)
                off += n
        self._flat = flat
        self._flat_grad = None
        self._packed_key = None
        self.__dict__["_pad"] = None
        self.__dict__["_native_grads"] = {}

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._release()
        self._flatten()  # .to()/.cuda() re-allocates every parameter: gather them into one buffer again
        return out

    def flat_params(self) -> torch.Tensor:
        self._check_views()
        return self._flat

    def flat_grad(self) -> torch.Tensor:
        if self._flat_grad is None or self._flat_grad.device != self._flat.device:
            self._flat_grad = torch.zeros_like(self._flat)
        return self._flat_grad

    def param_slices(self):
        """[(parameter, offset, numel)] in flat order."""
        out, off = [], 0
        for p in self._ordered_params():
            out.append((p, off, p.numel()))
            off += p.numel()
        return out

    def _check_views(self):
        base = self._flat.data_ptr()
        off = 0
        for p in self._ordered_params():
            if p.data_ptr() != base + 4 * off or p.device != self._flat.device:
                self._flatten()  # someone re-pointed a parameter (e.g. p.data = ...): re-gather
                return
            off += p.numel()

    # ---- native handle ------------------------------------------------------------------------------------
    _ABI = ("frl_net_create", "frl_net_destroy", "frl_net_set_params")

    def _native_config(self, device_index: int):
        return _native.NetConfig(self.s_dim, self.a_dim, self.native_hidden(), self.depth, self.n_heads, device_index)

    def _release(self):
        if getattr(self, "_handle", None):
            getattr(_native.lib(), self._ABI[1])(self._handle)
        self._handle = None
        self._packed_key = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def native(self) -> C.c_void_p:
        """Handle with up-to-date packed weights (re-packs iff a parameter changed since the last call)."""
        self._check_views()
        dev = self._flat.device
        if dev.type != "cuda":
            raise RuntimeError(
                "modril_b200 flow nets compute only on CUDA (sm_100a) devices; there is no CPU fallback. "
                "Move the module with .to('cuda').")
        L = _native.lib()
        if self._handle is None:
            cfg = self._native_config(dev.index or 0)
            h = C.c_void_p()
            _native.check(getattr(L, self._ABI[0])(C.byref(cfg), C.byref(h)), self._ABI[0])
            self._handle = h
        key = (self._flat.data_ptr(), tuple(p._version for p in self._ordered_params()))
        if key != self._packed_key:
            src = self._flat
            pad = self._pad_plan()
            if pad is not None:   # refresh the zero-padded image the library packs from
                src = pad[2]
                with torch.no_grad():
                    src.index_copy_(0, pad[0], self._flat.detach())
            _native.check(getattr(L, self._ABI[2])(self._handle, _ptr(src), _current_stream(dev)), self._ABI[2])
            self._packed_key = key
        return self._handle

    def train_timeouts(self) -> int:
        """Diagnostic: global waits of the fused tensor-core update that gave up on this net (0 when healthy).  Synchronises."""
        if self._handle is None or not hasattr(_native.lib(), "frl_train_timeouts") or self._ABI[0] != "frl_net_create":
            return 0
        return int(_native.lib().frl_train_timeouts(self._handle))

    def mark_dirty(self):
        """Force a re-pack on next use (needed only after writing weights through ``.data`` or raw pointers)."""
        self._packed_key = None

    def _vfield(self, a, s, t):
        dev = self._flat.device
        h = self.native()
        a = as_f32_cuda(a, dev, "a")
        s = as_f32_cuda(s, dev, "s")
        t = as_f32_cuda(t, dev, "t").reshape(-1)
        B = a.shape[0]
        if a.shape != (B, self.a_dim) or s.shape != (B, self.s_dim) or t.numel() != B:
            raise ValueError(f"expected a [B,{self.a_dim}], s [B,{self.s_dim}], t [B,1]; got {tuple(a.shape)}, "
                             f"{tuple(s.shape)}, {tuple(t.shape)}")
        out0 = torch.empty(B, self.a_dim, device=dev)
        out1 = torch.empty(B, self.a_dim, device=dev) if self.n_heads == 2 else None
        _native.check(_native.lib().frl_vfield(h, _ptr(s), _ptr(a), _ptr(t), _ptr(out0), _ptr(out1), B,
                                               _native.PREC_FP32, _current_stream(dev)), "frl_vfield")
        return out0, out1


class _VelocityNet(_FlatVNet):
    def native_hidden(self) -> int:
        if not 1 <= self.hidden <= 256:
            raise ValueError(f"hidden_dim {self.hidden}: the sm_100a kernels cover trunk widths up to 256")
        return 128 if self.hidden <= 128 else 256


class SharedVNet(_VelocityNet):
    """Shared trunk with two heads: returns (v_c, tanh(r))  -- reference networks.py:5-25."""

    n_heads = 2

    def __init__(self, s_dim, a_dim, hidden=256, depth=4):
        super().__init__()
        self.depth = max(1, depth)
        layers = [nn.Linear(s_dim + a_dim + 1, hidden), nn.ReLU()]
        for _ in range(self.depth - 1):
            layers += [nn.Linear(hidden, hidden), nn.ReLU()]
        self.shared = nn.Sequential(*layers)
        self.head_c = nn.Linear(hidden, a_dim)
        self.head_r = nn.Linear(hidden, a_dim)
        self._finish_init(s_dim, a_dim, hidden)

    def forward(self, a, s, t):
        return self._vfield(a, s, t)


class ConditionalVNet(_VelocityNet):
    """Time-conditioned vector field v(a_t, s, t) -- reference networks.py:28-49."""

    n_heads = 1

    def __init__(self, s_dim, a_dim, hidden=128, depth=4):
        super().__init__()
        self.depth = max(1, depth)
        layers = [nn.Linear(a_dim + s_dim + 1, hidden), nn.ReLU()]
        for _ in range(self.depth - 1):
            layers += [nn.Linear(hidden, hidden), nn.ReLU()]
        layers.append(nn.Linear(hidden, a_dim))
        self.net = nn.Sequential(*layers)
        self._finish_init(s_dim, a_dim, hidden)

    def forward(self, a_t, s, t):
        return self._vfield(a_t, s, t)[0]


class GradNorm2D(nn.Module):
    """GradNorm balancing of the two regulariser weights -- same contract as reference networks.py:52-138: constructor
    arguments, parameters (``log_beta``, ``log_gamma``), buffers (``step_count``, ``L0_anti``, ``L0_jpen``,
    ``ema_g_anti``, ``ema_g_jpen``) and update rule, so its ``state_dict`` round-trips with the reference's.

    Differences in mechanism: (1) the reference's ``update`` obtains the two gradient norms with
    ``torch.autograd.grad(loss, params)`` (networks.py:108-117); here the regulariser gradients come out of the fused
    kernels (``frl_reg_loss_grad``), so ``update`` takes the two sums of per-tensor norms directly (``g_anti=`` /
    ``g_jpen=``, see ``pretrain.tensor_grad_norm_sum``).  (2) The step counter has a host-side mirror (``steps``): the
    branches on it (warm-up, update cadence) never read the device buffer, so they cost no device->host sync.
    """

    _TERMS = (("anti", "beta"), ("jpen", "gamma"))   # (buffer suffix, weight name) of the two regularisers

    def __init__(self, beta_init=1e-4, gamma_init=1e-5, alpha=0.1, eps=1e-8, beta_min=1e-6, beta_max=1e1,
                 gamma_min=1e-8, gamma_max=1e-2, warmup_steps=100, update_freq=10, ema_decay=0.9, enable_beta=True,
                 enable_gamma=True):
        super().__init__()
        self.alpha, self.eps, self.ema_decay = alpha, eps, ema_decay
        self.warmup_steps, self.update_freq = warmup_steps, update_freq
        self.beta_min, self.beta_max, self.gamma_min, self.gamma_max = beta_min, beta_max, gamma_min, gamma_max
        self.enable_beta, self.enable_gamma = enable_beta, enable_gamma
        for name, init in (("beta", beta_init), ("gamma", gamma_init)):
            setattr(self, "log_" + name, nn.Parameter(torch.log(torch.tensor(init))))
        self.register_buffer("step_count", torch.tensor(0))
        for buf in ("L0_anti", "L0_jpen", "ema_g_anti", "ema_g_jpen"):
            self.register_buffer(buf, torch.tensor(0.))
        self._steps = 0

    @property
    def steps(self) -> int:
        """Host mirror of ``step_count`` (the number of ``forward`` calls so far)."""
        return self._steps

    def set_steps(self, n: int):
        """Set the step counter (device buffer and host mirror) -- what advancing ``step_count`` externally means here."""
        self._steps = int(n)
        self.step_count.fill_(int(n))

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        key = prefix + "step_count"
        if key in state_dict:
            self._steps = int(state_dict[key])

    def _enabled(self, weight):
        return self.enable_beta if weight == "beta" else self.enable_gamma

    def forward(self, loss_anti, j_pen):
        losses = {"anti": loss_anti, "jpen": j_pen}
        n = self._steps
        if n < self.warmup_steps:   # L0 = mean of the two losses over the first warmup_steps calls
            for suffix, _ in self._TERMS:
                ref = getattr(self, "L0_" + suffix)
                ref += losses[suffix].detach()
                if n + 1 == self.warmup_steps:
                    ref /= float(self.warmup_steps)
        self._steps = n + 1
        self.step_count += 1
        w = {name: (getattr(self, "log_" + name).exp() if self._enabled(name)
                    else torch.tensor(0.0, device=losses[suffix].device)) for suffix, name in self._TERMS}
        return w["beta"] * loss_anti + w["gamma"] * j_pen, w["beta"], w["gamma"]

    @torch.no_grad()
    def update(self, loss_anti, j_pen, params=None, *, g_anti=0.0, g_jpen=0.0):
        import math

        if self._steps <= self.warmup_steps or self._steps % self.update_freq != 0:
            return
        losses = {"anti": loss_anti, "jpen": j_pen}
        norms = {"anti": float(g_anti) if self.enable_beta else 0.0, "jpen": float(g_jpen) if self.enable_gamma else 0.0}
        g, r = {}, {}
        for suffix, _ in self._TERMS:   # smoothed gradient norms and loss ratios L / L0 of the two terms
            ema = self.ema_decay * getattr(self, "ema_g_" + suffix) + (1 - self.ema_decay) * norms[suffix]
            setattr(self, "ema_g_" + suffix, ema)
            g[suffix] = float(ema)
            r[suffix] = (losses[suffix].detach() / (getattr(self, "L0_" + suffix) + self.eps)).item()
        r_mean = 0.5 * (r["anti"] + r["jpen"]) + self.eps
        g_mean = 0.5 * (g["anti"] + g["jpen"]) + self.eps
        for suffix, name in self._TERMS:
            if not self._enabled(name):
                continue
            log_w = getattr(self, "log_" + name)
            log_w.data += self.alpha * (g[suffix] / g_mean * (r[suffix] / r_mean) - 1)
            log_w.data.clamp_(math.log(getattr(self, name + "_min")), math.log(getattr(self, name + "_max")))
