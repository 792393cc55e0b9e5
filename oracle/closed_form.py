"""CPU oracle for the FLOWRIL flow-matching density hot path (numpy, forward-mode restatement).

TEST INFRASTRUCTURE ONLY.  Nothing under ``modril_b200/`` may import this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use it, and there
only as the checker / the CPU arm -- never as the product path.

Parity status: the reference's own test-suite holds no golden vector for this path, so
the oracle is pinned against outputs of the reference code itself, run in the build container by
``oracle/gen_golden.py`` and committed under ``tests/golden/`` (see ``tests/test_oracle_vs_golden.py``).

Every function restates reference arithmetic and cites the reference ``file:line`` it follows (paths relative to
the reference checkout).  The restatement is *not* a transcription: the reference computes the divergence with a
reverse-mode VJP through autograd (``modril/flowril/pretrain.py:163-177``); here the same quantity
``v^T (dv/da) v`` is carried as a forward-mode tangent next to the primal, and gradients of the CFM loss are
written out by hand, so that the oracle shares no code path (autograd) with the reference.

Parameters travel as one flat vector in ``state_dict`` order:

  scrf  (SharedVNet,      modril/flowril/networks.py:10-21):
        shared.0.weight [H, a+s+1], shared.0.bias [H], (depth-1) x {shared.2l.weight [H,H], .bias [H]},
        head_c.weight [a,H], head_c.bias [a], head_r.weight [a,H], head_r.bias [a]
  2fs   (ConditionalVNet, modril/flowril/networks.py:34-44):
        net.0.weight [H, a+s+1], net.0.bias [H], (depth-1) x {...}, net.2*depth.weight [a,H], .bias [a]

Input column order of the first layer is ``[a | s | t]`` (networks.py:24, :48).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

LOG_SQRT_2PI = 0.5 * math.log(2.0 * math.pi)


@dataclass(frozen=True)
class NetSpec:
    """Shape of one velocity net.  ``n_heads`` = 2 for the coupled-residual SharedVNet, 1 for ConditionalVNet."""

    s_dim: int
    a_dim: int
    hidden: int = 256
    depth: int = 4
    n_heads: int = 2

    @property
    def d_in(self) -> int:
        return self.a_dim + self.s_dim + 1

    def shapes(self):
        """[(name, shape)] in state_dict order (networks.py:13-21, :37-44)."""
        depth = max(1, self.depth)  # networks.py:12, :36
        out = []
        prefix = "net.shared" if self.n_heads == 2 else "net.net"
        k = self.d_in
        for layer in range(depth):
            out.append((f"{prefix}.{2 * layer}.weight", (self.hidden, k)))
            out.append((f"{prefix}.{2 * layer}.bias", (self.hidden,)))
            k = self.hidden
        if self.n_heads == 2:
            for head in ("head_c", "head_r"):
                out.append((f"net.{head}.weight", (self.a_dim, self.hidden)))
                out.append((f"net.{head}.bias", (self.a_dim,)))
        else:
            out.append((f"{prefix}.{2 * depth}.weight", (self.a_dim, self.hidden)))
            out.append((f"{prefix}.{2 * depth}.bias", (self.a_dim,)))
        return out

    def n_params(self) -> int:
        return sum(int(np.prod(s)) for _, s in self.shapes())


def unpack(spec: NetSpec, flat: np.ndarray):
    """flat vector -> dict name -> array view."""
    out, off = {}, 0
    for name, shape in spec.shapes():
        n = int(np.prod(shape))
        out[name] = flat[off:off + n].reshape(shape)
        off += n
    assert off == flat.size, (off, flat.size)
    return out


def make_params(spec: NetSpec, seed: int, gain: float = 1.0, dtype=np.float32) -> np.ndarray:
    """Deterministic synthetic weights: U(+-gain/sqrt(fan_in)) like nn.Linear's default bound
    (the reference never overrides the torch default init, networks.py:14-21).  Numpy PCG64 is
    version-stable, so goldens store only (seed, gain), not the weights."""
    rng = np.random.Generator(np.random.PCG64(seed))
    parts = []
    for name, shape in spec.shapes():
        fan_in = shape[1] if len(shape) == 2 else None
        if fan_in is None:  # bias: bound from the preceding weight's fan_in
            fan_in = last_fan_in
        last_fan_in = fan_in
        bound = gain / math.sqrt(fan_in)
        parts.append(rng.uniform(-bound, bound, size=shape).astype(dtype).ravel())
    return np.concatenate(parts)


def _layers(spec: NetSpec, flat: np.ndarray):
    p = unpack(spec, flat)
    names = [n for n, _ in spec.shapes()]
    depth = max(1, spec.depth)
    trunk = [(p[names[2 * l]], p[names[2 * l + 1]]) for l in range(depth)]
    heads = [(p[names[2 * (depth + h)]], p[names[2 * (depth + h) + 1]]) for h in range(spec.n_heads)]
    return trunk, heads


# --------------------------------------------------------------------------------------------------
# a1/a2/a3: velocity field  (networks.py:23-25, :46-49; pretrain.py:156-158)
# --------------------------------------------------------------------------------------------------
def net_forward(spec: NetSpec, flat, a, s, t, tangent=None):
    """Returns (outs, douts, acts): outs = list of head pre-activations [B,a] (2 for scrf, 1 for 2fs);
    douts = same for the tangent (None if tangent is None); acts = per-layer (input, mask) for backprop.

    tangent: [B, a_dim] direction in action space.  Only the first a_dim columns of W_1 touch it because
    the input is cat[a, s, t] (networks.py:24) and s, t do not depend on a.
    """
    trunk, heads = _layers(spec, flat)
    x = np.concatenate([a, s, t], axis=1)
    tau = None
    acts = []
    for li, (W, b) in enumerate(trunk):
        z = x @ W.T + b
        m = z > 0  # nn.ReLU: relu'(0) = 0 (threshold_backward), so the mask is strict
        acts.append((x, m))
        if tangent is not None:
            zt = (tangent @ W[:, :spec.a_dim].T) if li == 0 else (tau @ W.T)
            tau = np.where(m, zt, 0.0)
        x = np.where(m, z, 0.0)
    outs = [x @ W.T + b for (W, b) in heads]
    douts = [tau @ W.T for (W, b) in heads] if tangent is not None else None
    acts.append((x, None))
    return outs, douts, acts


def v_field(spec: NetSpec, flat, a, s, t, role: str, tangent=None):
    """v_E = v_c + tanh(r), v_A = v_c - tanh(r) for scrf (pretrain.py:156-158; networks.py:25);
    plain head for 2fs (networks.py:46-49).  Returns (v, dv) with dv the directional derivative along
    `tangent` (None if not requested)."""
    outs, douts, _ = net_forward(spec, flat, a, s, t, tangent)
    if spec.n_heads == 1:
        return outs[0], (douts[0] if douts is not None else None)
    sign = 1.0 if role == "expert" else -1.0  # pretrain.py:158: anything but "expert" is the agent
    r = np.tanh(outs[1])
    v = outs[0] + sign * r
    dv = None
    if douts is not None:
        dv = douts[0] + sign * (1.0 - r * r) * douts[1]
    return v, dv


def torch_linspace_f32(n_steps: int) -> np.ndarray:
    """torch.linspace(0, 1, n+1) in float32 as torch's CPU kernel computes it: the step is a float32, the first half
    is start + step*i, the second half end - step*(steps-1-i), each evaluated as ONE fused multiply-add (a single
    rounding; emulated here in float64, which is exact for these operands).  The reference calls it at
    pretrain.py:114, :200.  Checked bit-for-bit against torch in tests/test_host_logic.py.  (The product computes
    its grid with torch on the model's device, like the reference does.)"""
    steps = n_steps + 1
    step = float(np.float32(1.0) / np.float32(steps - 1))
    half = steps // 2
    out = np.empty(steps, dtype=np.float32)
    for i in range(steps):
        out[i] = np.float32(0.0 + step * i) if i < half else np.float32(1.0 - step * (steps - i - 1))
    return out


def t_mid_grid(n_steps: int) -> np.ndarray:
    """t_mid_k = 0.5 * (g_k + g_{k+1}) in float32  (pretrain.py:122, :206)."""
    g = torch_linspace_f32(n_steps)
    return (np.float32(0.5) * (g[:-1] + g[1:])).astype(np.float32)


# --------------------------------------------------------------------------------------------------
# a4/a5: log-density  (pretrain.py:112-134, :192-214) with the divergence of pretrain.py:88-92, :163-177
# --------------------------------------------------------------------------------------------------
def log_prob(spec: NetSpec, flat, s, a0, role: str, probe, n_steps: int = 32, dtype=np.float64,
             exact: bool = False):
    """log p(a|s) = sum_j log N(a_1j; 0, 1) - sum_k delta * div_k, explicit Euler with midpoint time.

    probe: [B, a_dim] (same probe at every step -- scrf semantics) or [n_steps, B, a_dim]
    (fresh probe per step -- 2fs semantics, pretrain.py:90).  div = sum_j (J probe)_j probe_j = probe^T J probe,
    identical to the reference's (J^T probe) . probe.  exact=True sums e_i^T J e_i over the a_dim basis vectors
    (the trace itself; coincides with Hutchinson when a_dim == 1).
    """
    flat = np.asarray(flat, dtype=dtype)
    s = np.asarray(s, dtype=dtype)
    a = np.asarray(a0, dtype=dtype).copy()
    B = a.shape[0]
    tm = t_mid_grid(n_steps).astype(dtype)
    delta = dtype(1.0 / n_steps)  # pretrain.py:115, :201 (python float folded into the tensor dtype)
    log_det = np.zeros(B, dtype=dtype)
    for k in range(n_steps):
        t = np.full((B, 1), tm[k], dtype=dtype)
        if exact:
            div = np.zeros(B, dtype=dtype)
            v = None
            for i in range(spec.a_dim):
                e = np.zeros((B, spec.a_dim), dtype=dtype)
                e[:, i] = 1.0
                v, dv = v_field(spec, flat, a, s, t, role, tangent=e)
                div += dv[:, i]
        else:
            pk = np.asarray(probe[k] if np.ndim(probe) == 3 else probe, dtype=dtype)
            v, dv = v_field(spec, flat, a, s, t, role, tangent=pk)
            div = (dv * pk).sum(axis=1)
        log_det = log_det - delta * div  # pretrain.py:130, :210
        a = a + v * delta  # pretrain.py:131, :211
    logp_prior = (-0.5 * a * a - dtype(LOG_SQRT_2PI)).sum(axis=1)  # Normal(0,1).log_prob, pretrain.py:133, :213
    return logp_prior + log_det


# --------------------------------------------------------------------------------------------------
# a9: reward transform  (flowril.py:191-207)
# --------------------------------------------------------------------------------------------------
REWARD_TYPES = ("raw", "norm", "exp", "clip", "smooth")


def reward_transform(r, reward_type: str):
    r = np.asarray(r)
    one = r.dtype.type(1.0)
    if reward_type == "raw":
        return r
    if reward_type == "norm":
        sg = one / (one + np.exp(-r))
        eps = r.dtype.type(1e-20)
        return np.log(sg + eps) - np.log(one - sg + eps)
    if reward_type == "exp":
        e = np.exp(r)
        return -e / (one + e)
    if reward_type == "clip":
        return np.clip(r, -20.0, 20.0).astype(r.dtype)
    if reward_type == "smooth":
        return -np.log1p(np.exp(-r))
    raise ValueError(f"Unrecognized reward type {reward_type}")  # flowril.py:207


# --------------------------------------------------------------------------------------------------
# a6: CFM loss + hand-written gradient  (pretrain.py:94-109, :179-190)
# --------------------------------------------------------------------------------------------------
def fm_loss_and_grad(spec: NetSpec, flat, s, a0, role: str, t, noise, eps_unused=None, dtype=np.float64,
                     dequant=None):
    """L = mean_b (1-t_b)^1.5 * sum_j (v_role(a_t, s, t) - (noise - a0))_bj^2 with a_t = (1-t) a0 + t noise.

    t [B,1] and noise [B,a] are caller-supplied (the reference draws them at pretrain.py:182-183 / :99-101).
    dequant [B,a] (2fs only) is the already-scaled dequantisation noise added to a0 first (pretrain.py:98).
    Returns (loss, grad_flat) with grad in the same flat layout as the parameters.
    """
    flat = np.asarray(flat, dtype=dtype)
    s = np.asarray(s, dtype=dtype)
    a0 = np.asarray(a0, dtype=dtype)
    t = np.asarray(t, dtype=dtype).reshape(-1, 1)
    noise = np.asarray(noise, dtype=dtype)
    if dequant is not None:
        a0 = a0 + np.asarray(dequant, dtype=dtype)
    B = a0.shape[0]
    a_t = (1.0 - t) * a0 + t * noise
    v_star = noise - a0  # pretrain.py:85-86, :160-161
    outs, _, acts = net_forward(spec, flat, a_t, s, t)
    sign = 1.0 if role == "expert" else -1.0
    if spec.n_heads == 2:
        r = np.tanh(outs[1])
        v = outs[0] + sign * r
    else:
        v = outs[0]
    w = np.power(1.0 - t, dtype(1.5))  # pretrain.py:105, :188
    diff = v - v_star
    loss = float((w * (diff * diff).sum(axis=1, keepdims=True)).mean())

    # ---- backward ----
    dv = (2.0 / B) * w * diff  # [B, a]
    trunk, heads = _layers(spec, flat)
    h_last = acts[-1][0]
    grads = {}
    names = [n for n, _ in spec.shapes()]
    depth = max(1, spec.depth)
    if spec.n_heads == 2:
        d_c = dv
        d_r = dv * sign * (1.0 - r * r)
        douts = [d_c, d_r]
    else:
        douts = [dv]
    dh = np.zeros_like(h_last)
    for hidx, (W, b) in enumerate(heads):
        grads[names[2 * (depth + hidx)]] = douts[hidx].T @ h_last
        grads[names[2 * (depth + hidx) + 1]] = douts[hidx].sum(axis=0)
        dh = dh + douts[hidx] @ W
    for li in reversed(range(depth)):
        W, b = trunk[li]
        x_in, m = acts[li]
        dz = np.where(m, dh, 0.0)
        grads[names[2 * li]] = dz.T @ x_in
        grads[names[2 * li + 1]] = dz.sum(axis=0)
        if li > 0:
            dh = dz @ W
    gflat = np.concatenate([grads[n].ravel() for n in names]).astype(dtype)
    return loss, gflat


# --------------------------------------------------------------------------------------------------
# f1: anti-divergence and Jacobian-stability regularisers  (flowril.py:255-264; pretrain.py:55-59, :163-177)
# --------------------------------------------------------------------------------------------------
def reg_losses_and_grads(spec: NetSpec, flat, s, a, t, probe, dtype=np.float64):
    """The two regularisers of the scrf reward model on the mixed (expert + agent) batch and their gradients.

      loss_anti   = mean_b ( v^T J_r v )^2            J_r = d tanh(head_r) / d a     (flowril.py:262, _hutch_div k=1)
      loss_stable = mean_b || J_{v_c + r}^T v ||^2                                   (flowril.py:263, pretrain.py:55-59)

    with v = probe [N, a] (the reference draws the same shape-deterministic `_rademacher` matrix for both),
    evaluated at (a, s, t) with t = torch.rand(N, 1) (flowril.py:259 -- no eps rescale).

    The reference differentiates these through autograd's double backward.  Restated by hand (ReLU has zero second
    derivative a.e., so only tanh'' and the weights' bilinear terms survive):

      anti:    forward tangent u_l = D_l W_l u_{l-1} (u_0 = v on the action columns), w = W_r u_L,
               q = sum_j v_j tau_j w_j with tau = 1 - r^2; adjoints: g_w = dq * v * tau back through the tangent
               chain (weight grads = outer products with u_{l-1}); g_rho = dq * v * w * (-2 r tau) back through the
               primal chain (ordinary MLP backward).
      stable:  reverse chain mu_L = W_c^T v + W_r^T (v * tau), nu_l = D_l mu_l, mu_{l-1} = W_l^T nu_l,
               g = A_1^T nu_1 (A_1 = action columns of W_1), p = |g|^2; its adjoint is a FORWARD tangent pass
               mubar_0 = dp * 2 g, nubar_l = W_l mubar_{l-1}, mubar_l = D_l nubar_l with weight grads
               nu_l (outer) mubar_{l-1}; heads: dW_c += v (outer) mubar_L, dW_r += (v tau) (outer) mubar_L,
               g_rho = v * (W_r mubar_L) * (-2 r tau) back through the primal chain.

    Returns (loss_anti, loss_stable, grad_anti_flat, grad_stable_flat).  Requires n_heads == 2.
    """
    assert spec.n_heads == 2, "the regularisers exist only for the coupled-residual (scrf) model"
    flat = np.asarray(flat, dtype=dtype)
    s = np.asarray(s, dtype=dtype)
    a = np.asarray(a, dtype=dtype)
    t = np.asarray(t, dtype=dtype).reshape(-1, 1)
    v = np.asarray(probe, dtype=dtype)
    N, A = a.shape
    depth = max(1, spec.depth)
    trunk, heads = _layers(spec, flat)
    (Wc, bc), (Wr, br) = heads
    names = [n for n, _ in spec.shapes()]

    # primal + tangent forward
    x = np.concatenate([a, s, t], axis=1)
    hs, us, masks = [x], [None], []
    u = None
    for li, (W, b) in enumerate(trunk):
        z = hs[-1] @ W.T + b
        m = z > 0
        zt = (v @ W[:, :A].T) if li == 0 else (u @ W.T)
        u = np.where(m, zt, 0.0)
        masks.append(m)
        hs.append(np.where(m, z, 0.0))
        us.append(u)
    hL, uL = hs[-1], us[-1]
    r = np.tanh(hL @ Wr.T + br)
    tau = 1.0 - r * r
    dtau_drho = -2.0 * r * tau

    def primal_backward(g_rho, grads):
        """ordinary MLP backward from a cotangent on the head_r pre-activation"""
        grads[names[2 * depth + 2]] += g_rho.T @ hL
        grads[names[2 * depth + 3]] += g_rho.sum(axis=0)
        dh = g_rho @ Wr
        for li in reversed(range(depth)):
            dz = np.where(masks[li], dh, 0.0)
            grads[names[2 * li]] += dz.T @ hs[li]
            grads[names[2 * li + 1]] += dz.sum(axis=0)
            if li > 0:
                dh = dz @ trunk[li][0]

    def zero_grads():
        return {n: np.zeros(shape, dtype=dtype) for n, shape in spec.shapes()}

    # ---- anti ----
    w = uL @ Wr.T
    q = (v * tau * w).sum(axis=1)
    loss_anti = float((q * q).mean())
    ga = zero_grads()
    dq = (2.0 / N) * q[:, None]
    g_w = dq * v * tau
    ga[names[2 * depth + 2]] += g_w.T @ uL
    du = g_w @ Wr
    for li in reversed(range(depth)):
        dzt = np.where(masks[li], du, 0.0)
        if li == 0:
            ga[names[0]][:, :A] += dzt.T @ v
        else:
            ga[names[2 * li]] += dzt.T @ us[li]
            du = dzt @ trunk[li][0]
    primal_backward(dq * v * w * dtau_drho, ga)

    # ---- stable ----
    mu = v @ Wc + (v * tau) @ Wr
    nus = [None] * depth
    for li in reversed(range(depth)):
        nus[li] = np.where(masks[li], mu, 0.0)
        if li > 0:
            mu = nus[li] @ trunk[li][0]
    g = nus[0] @ trunk[0][0][:, :A]
    loss_stable = float((g * g).sum(axis=1).mean())
    gs = zero_grads()
    mubar = (2.0 / N) * g                       # [N, A]
    gs[names[0]][:, :A] += nus[0].T @ mubar
    nubar = mubar @ trunk[0][0][:, :A].T
    mubar = np.where(masks[0], nubar, 0.0)
    for li in range(1, depth):
        gs[names[2 * li]] += nus[li].T @ mubar
        nubar = mubar @ trunk[li][0].T
        mubar = np.where(masks[li], nubar, 0.0)
    gs[names[2 * depth]] += v.T @ mubar
    gs[names[2 * depth + 2]] += (v * tau).T @ mubar
    primal_backward(v * (mubar @ Wr.T) * dtau_drho, gs)

    cat = lambda d: np.concatenate([d[n].ravel() for n in names]).astype(dtype)
    return loss_anti, loss_stable, cat(ga), cat(gs)


# --------------------------------------------------------------------------------------------------
# a8: clip_grad_norm_ + Adam  (flowril.py:324-330; torch.nn.utils.clip_grad_norm_, torch.optim.Adam defaults)
# --------------------------------------------------------------------------------------------------
def clip_coef(grad, max_norm: float, trainable=None):
    """torch.nn.utils.clip_grad_norm_: total 2-norm over params that HAVE a grad (frozen params have
    grad None and are skipped -- SURVEY 7.2-5); coef = max_norm / (norm + 1e-6) clamped to <= 1."""
    g = grad if trainable is None else grad[trainable]
    total = math.sqrt(float((g.astype(np.float64) ** 2).sum()))
    return min(1.0, max_norm / (total + 1e-6)), total


def adam_step(flat, grad, m, v, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8, trainable=None):
    """One torch.optim.Adam update (defaults betas=(0.9,0.999), eps=1e-8, weight_decay=0, amsgrad=False;
    the reference constructs it at flowril.py:60-65, :106, :109).  `step` is the 1-based step count AFTER
    increment.  Returns new (flat, m, v); entries where trainable is False are untouched."""
    dtype = flat.dtype
    m_new = beta1 * m + (1.0 - beta1) * grad
    v_new = beta2 * v + (1.0 - beta2) * grad * grad
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = lr / bc1
    denom = np.sqrt(v_new) / math.sqrt(bc2) + eps
    p_new = flat - step_size * (m_new / denom)
    if trainable is not None:
        p_new = np.where(trainable, p_new, flat)
        m_new = np.where(trainable, m_new, m)
        v_new = np.where(trainable, v_new, v)
    return p_new.astype(dtype), m_new.astype(dtype), v_new.astype(dtype)


def train_step(spec: NetSpec, flat, m, v, step, batches, lr, max_norm, rates=(1.0, 1.0), trainable=None,
               dtype=np.float64, reg=None):
    """One reward-model update as flowril.py:311-335 does it for scrf:
    loss = rate_E * L(role expert batch) + rate_A * L(role agent batch) [+ w_anti * L_anti + w_stable * L_stable];
    clip; Adam.  batches = [(role, s, a0, t, noise), ...] in the order the losses are summed.

    reg (flowril.py:255-284): dict(s, a, t, probe, anti=spec, stable=spec) on the mixed batch, spec = None (off),
    a float (fixed weight) or ("warmup", scale, eps) for w = L_E / (L_reg + eps) * scale with L_E the first loss.
    With reg the return value gains (loss_anti, loss_stable, w_anti, w_stable) as a fifth element."""
    total = np.zeros_like(flat, dtype=dtype)
    losses = []
    for (role, s, a0, t, noise), rate in zip(batches, rates):
        l, g = fm_loss_and_grad(spec, flat, s, a0, role, t, noise, dtype=dtype)
        losses.append(l)
        total += rate * g
    reg_out = None
    if reg is not None:
        la, ls, ga, gs = reg_losses_and_grads(spec, flat, reg["s"], reg["a"], reg["t"], reg["probe"], dtype=dtype)
        ws = []
        for sp, l_reg, g_reg in ((reg.get("anti"), la, ga), (reg.get("stable"), ls, gs)):
            if sp is None:
                w = 0.0
            elif isinstance(sp, tuple):
                w = losses[0] / (l_reg + sp[2]) * sp[1]
            else:
                w = float(sp)
            total += w * g_reg
            ws.append(w)
        reg_out = (la, ls, ws[0], ws[1])
    if trainable is not None:
        total = np.where(trainable, total, 0.0)
    coef, norm = clip_coef(total, max_norm, trainable)
    total = total * coef
    flat, m, v = adam_step(np.asarray(flat, dtype=dtype), total, m, v, step, lr, trainable=trainable)
    if reg is not None:
        return flat, m, v, losses, norm, reg_out
    return flat, m, v, losses, norm


# --------------------------------------------------------------------------------------------------
# a10: return normalisation  (flowril.py:221-228; rlf/baselines/common/running_mean_std.py:3-35)
# --------------------------------------------------------------------------------------------------
class RunningMeanStd:
    """float32 running moments with the parallel-variance merge (running_mean_std.py:3-35)."""

    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape, "float32")
        self.var = np.ones(shape, "float32")
        self.count = epsilon

    def update(self, x):
        bm = np.mean(x, axis=0, dtype="float32")
        bv = np.var(x, axis=0, dtype="float32")
        n = x.shape[0]
        delta = bm - self.mean
        tot = self.count + n
        new_mean = self.mean + delta * n / tot
        m2 = self.var * self.count + bv * n + np.square(delta, dtype="float32") * self.count * n / tot
        self.mean, self.var, self.count = new_mean, m2 / tot, tot


def normalise_rewards(rewards, masks, gamma: float, ret_rms: RunningMeanStd | None = None, returns=None):
    """Sequential scan over rollout steps (flowril.py:221-228): rewards [T,N,1] float32, masks [T,N,1].
    First call: returns = reward.clone() and then the recursion is *still* applied (flowril.py:222-225).
    Returns (normalised rewards [T,N,1], final returns, ret_rms)."""
    ret_rms = ret_rms or RunningMeanStd(shape=())
    out = np.empty_like(rewards, dtype=np.float32)
    for step in range(rewards.shape[0]):
        r = rewards[step].astype(np.float32)
        if returns is None:
            returns = r.copy()
        returns = (returns * masks[step].astype(np.float32) * np.float32(gamma) + r).astype(np.float32)
        ret_rms.update(returns)
        out[step] = r / np.sqrt(ret_rms.var[0] + 1e-8)
    return out, returns, ret_rms
