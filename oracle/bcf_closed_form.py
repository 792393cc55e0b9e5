"""CPU oracle for the BCF flow-matching policy (scope row f3 of SURVEY.md section 8): numpy restatement of
``VectorNetwork`` (deps/rl-toolkit/rlf/rl/model.py:392-432), the 32-step Euler action sampler of ``FlowPolicy.get_action``
(rlf/policies/flow_policy.py:49-63) and the loss of ``BehaviorCloneFlowMatching._bc_step`` (rlf/algos/il/bcf.py:108-128)
with a hand-written backward.

TEST INFRASTRUCTURE ONLY (see oracle/closed_form.py): nothing under ``modril_b200/`` imports this module.  Pinned to
outputs of the reference's own ``VectorNetwork`` class by ``oracle/gen_golden.py`` (``tests/golden/bcf_*.npz``).

Flat parameter order = ``state_dict`` order of VectorNetwork:
  time_embed.0.weight [T,1], .bias [T], time_embed.2.weight [H,T], .bias [H], state_embed.0.weight [H,S], .bias [H],
  action_embed.0.weight [H,A], .bias [H], velocity_head.0.weight [H,3H], .bias [H], velocity_head.2.weight [A,H], .bias [A]
The head's input is cat[s_feat, t_feat, x_feat] (model.py:431).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class VNetSpec:
    s_dim: int
    a_dim: int
    hidden: int = 256
    time_dim: int = 128

    def shapes(self):
        H, T, S, A = self.hidden, self.time_dim, self.s_dim, self.a_dim
        return [("time_embed.0.weight", (T, 1)), ("time_embed.0.bias", (T,)),
                ("time_embed.2.weight", (H, T)), ("time_embed.2.bias", (H,)),
                ("state_embed.0.weight", (H, S)), ("state_embed.0.bias", (H,)),
                ("action_embed.0.weight", (H, A)), ("action_embed.0.bias", (H,)),
                ("velocity_head.0.weight", (H, 3 * H)), ("velocity_head.0.bias", (H,)),
                ("velocity_head.2.weight", (A, H)), ("velocity_head.2.bias", (A,))]

    def n_params(self):
        return sum(int(np.prod(s)) for _, s in self.shapes())


def unpack(spec: VNetSpec, flat):
    out, off = {}, 0
    for name, shape in spec.shapes():
        n = int(np.prod(shape))
        out[name] = flat[off:off + n].reshape(shape)
        off += n
    assert off == flat.size
    return out


def make_params(spec: VNetSpec, seed: int, gain: float = 1.0, dtype=np.float32):
    """Deterministic synthetic weights U(+-gain/sqrt(fan_in)) with NON-zero biases (the reference initialises
    orthogonal weights and zero biases, model.py:405-420; trained nets have neither, so the fixtures use generic values)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    parts, fan = [], 1
    for name, shape in spec.shapes():
        if len(shape) == 2:
            fan = shape[1]
        b = gain / np.sqrt(fan)
        parts.append(rng.uniform(-b, b, size=shape).astype(dtype).ravel())
    return np.concatenate(parts)


def forward(spec: VNetSpec, flat, x, s, t, keep=False):
    """VectorNetwork.forward (model.py:422-432): v = head(cat[relu(Ws s), relu(Wt2 relu(Wt1 t)), relu(Wa x)])."""
    p = unpack(spec, np.asarray(flat))
    t = np.asarray(t).reshape(-1, 1)
    z_t1 = t @ p["time_embed.0.weight"].T + p["time_embed.0.bias"]
    t1 = np.maximum(z_t1, 0)
    z_t2 = t1 @ p["time_embed.2.weight"].T + p["time_embed.2.bias"]
    t2 = np.maximum(z_t2, 0)
    z_s = s @ p["state_embed.0.weight"].T + p["state_embed.0.bias"]
    s1 = np.maximum(z_s, 0)
    z_x = x @ p["action_embed.0.weight"].T + p["action_embed.0.bias"]
    x1 = np.maximum(z_x, 0)
    hc = np.concatenate([s1, t2, x1], axis=1)
    z_h = hc @ p["velocity_head.0.weight"].T + p["velocity_head.0.bias"]
    h1 = np.maximum(z_h, 0)
    v = h1 @ p["velocity_head.2.weight"].T + p["velocity_head.2.bias"]
    if keep:
        return v, dict(t=t, t1=t1, t2=t2, s1=s1, x1=x1, hc=hc, h1=h1, m_t1=z_t1 > 0, m_t2=z_t2 > 0, m_s=z_s > 0,
                       m_x=z_x > 0, m_h=z_h > 0, s=s, x=x)
    return v


def backward(spec: VNetSpec, flat, cache, dv):
    """Gradient of sum(dv * v) w.r.t. the flat parameters."""
    p = unpack(spec, np.asarray(flat))
    H = spec.hidden
    g = {}
    g["velocity_head.2.weight"] = dv.T @ cache["h1"]
    g["velocity_head.2.bias"] = dv.sum(0)
    dzh = np.where(cache["m_h"], dv @ p["velocity_head.2.weight"], 0.0)
    g["velocity_head.0.weight"] = dzh.T @ cache["hc"]
    g["velocity_head.0.bias"] = dzh.sum(0)
    dhc = dzh @ p["velocity_head.0.weight"]
    dzs = np.where(cache["m_s"], dhc[:, :H], 0.0)
    dzt2 = np.where(cache["m_t2"], dhc[:, H:2 * H], 0.0)
    dzx = np.where(cache["m_x"], dhc[:, 2 * H:], 0.0)
    g["state_embed.0.weight"] = dzs.T @ cache["s"]
    g["state_embed.0.bias"] = dzs.sum(0)
    g["action_embed.0.weight"] = dzx.T @ cache["x"]
    g["action_embed.0.bias"] = dzx.sum(0)
    g["time_embed.2.weight"] = dzt2.T @ cache["t1"]
    g["time_embed.2.bias"] = dzt2.sum(0)
    dzt1 = np.where(cache["m_t1"], dzt2 @ p["time_embed.2.weight"], 0.0)
    g["time_embed.0.weight"] = dzt1.T @ cache["t"]
    g["time_embed.0.bias"] = dzt1.sum(0)
    return np.concatenate([g[n].ravel() for n, _ in spec.shapes()])


def bcf_loss_and_grad(spec: VNetSpec, flat, s, actions, x0, t, bc_coef=1.0, dtype=np.float64):
    """bcf.py:108-128: flow loss mse(v(x_t, s, t), a - x0) with x_t = (1-t) x0 + t a, plus the t = 1 behaviour-cloning
    term mse(x0 + v(a, s, 1), a); F.mse_loss means over all B*A elements.  x0 ~ N(0,I) and t ~ U[0,1) are caller-supplied.
    Returns (total, flow_loss, action_loss, grad_flat)."""
    flat = np.asarray(flat, dtype)
    s, a, x0 = (np.asarray(v, dtype) for v in (s, actions, x0))
    t = np.asarray(t, dtype).reshape(-1, 1)
    B, A = a.shape
    x_t = (1 - t) * x0 + t * a
    v1, c1 = forward(spec, flat, x_t, s, t, keep=True)
    d1 = v1 - (a - x0)
    flow = float((d1 * d1).mean())
    v2, c2 = forward(spec, flat, a, s, np.ones_like(t), keep=True)
    d2 = x0 + v2 - a
    act = float((d2 * d2).mean())
    grad = backward(spec, flat, c1, 2.0 / (B * A) * d1) + backward(spec, flat, c2, bc_coef * 2.0 / (B * A) * d2)
    return flow + bc_coef * act, flow, act, grad


def sample_actions(spec: VNetSpec, flat, s, x0, n_steps=32, dtype=np.float64):
    """FlowPolicy.get_action (flow_policy.py:49-63): x <- x0; for i < n: x += v(x, s, i*dt) * dt  (t is the LEFT end of the
    step, unlike the midpoint rule of the density model)."""
    flat = np.asarray(flat, dtype)
    s = np.asarray(s, dtype)
    x = np.asarray(x0, dtype).copy()
    dt = dtype(1.0 / n_steps)
    for i in range(n_steps):
        t = np.full((x.shape[0], 1), dtype(np.float32(i * (1.0 / n_steps))), dtype)   # python double i*dt -> fp32 tensor
        x = x + forward(spec, flat, x, s, t) * dt
    return x


def bc_step(spec: VNetSpec, flat, m, v, step, s, actions, x0, t, lr, bc_coef=1.0, max_norm=0.5, eps=1e-12,
            dtype=np.float64):
    """One ``_bc_step`` update: loss (bcf.py:108-128) then ``_standard_step`` (rlf/algos/base_net_algo.py:84-100):
    zero_grad, backward, clip_grad_norm_(max_grad_norm, default 0.5) and Adam.step (eps default 1e-12,
    base_net_algo.py:130-146).  Returns (flat, m, v, (loss, flow, act), pre-clip grad norm)."""
    from oracle import closed_form as cf

    total, flow, act, grad = bcf_loss_and_grad(spec, flat, s, actions, x0, t, bc_coef, dtype)
    norm = 0.0
    if max_norm > 0:
        coef, norm = cf.clip_coef(grad, max_norm)
        grad = grad * coef
    flat, m, v = cf.adam_step(np.asarray(flat, dtype), grad, m, v, step, lr, eps=eps)
    return flat, m, v, (total, flow, act), norm
