"""PyTorch-CPU port of the reference's hot path, used ONLY as the timed CPU baseline (bench.py `cpu_baseline` and
`--impl reference`) and cross-checked against the numpy oracle in tests/test_oracle_host.py.

TEST / BENCH INFRASTRUCTURE (see oracle/closed_form.py header) -- never imported by modril_b200.

Why a port: the reference is pure Python (`/root/reference`) and does not travel to the GPU box, so its own classes
cannot be timed there.  This port issues the same torch op sequence per Euler step as the reference does
(cat -> Linear/ReLU stack -> heads/tanh -> ONE reverse-mode VJP with create_graph=True -> elementwise updates;
modril/flowril/pretrain.py:192-214, :163-177; networks.py:23-25), so its CPU time is representative of the
reference's; it is written functionally over the flat parameter vector instead of nn.Modules.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import closed_form as cf


def _views(spec: cf.NetSpec, flat: torch.Tensor):
    out, off = [], 0
    for _, shape in spec.shapes():
        n = 1
        for d in shape:
            n *= d
        out.append(flat[off:off + n].view(shape))
        off += n
    return out


def velocity(spec: cf.NetSpec, views, a, s, t, sign: float):
    depth = max(1, spec.depth)
    h = torch.cat([a, s, t], dim=1)  # networks.py:24
    for l in range(depth):
        h = F.relu(F.linear(h, views[2 * l], views[2 * l + 1]))
    if spec.n_heads == 2:
        v_c = F.linear(h, views[2 * depth], views[2 * depth + 1])
        r = torch.tanh(F.linear(h, views[2 * depth + 2], views[2 * depth + 3]))  # networks.py:25
        return v_c + sign * r  # pretrain.py:156-158
    return F.linear(h, views[2 * depth], views[2 * depth + 1])


def log_prob(spec: cf.NetSpec, flat: torch.Tensor, s, a0, role: str, probe, n_steps: int = 32):
    """pretrain.py:192-214 with the probe injected (shared [B,a] or per-step [n,B,a])."""
    views = _views(spec, flat)
    sign = 1.0 if role == "expert" else -1.0
    B = a0.shape[0]
    grid = torch.linspace(0.0, 1.0, n_steps + 1)
    delta = 1.0 / n_steps
    log_det = torch.zeros(B)
    a = a0.detach()
    for k in range(n_steps):
        t_mid = 0.5 * (grid[k] + grid[k + 1])
        with torch.enable_grad():
            a = a.detach().requires_grad_(True)
            v = velocity(spec, views, a, s, t_mid.expand(B, 1), sign)
            pk = probe[k] if probe.dim() == 3 else probe
            (jv,) = torch.autograd.grad(v, a, pk, create_graph=True, retain_graph=True)  # pretrain.py:169-175
            div = (jv * pk).sum(dim=-1)
        log_det = log_det - delta * div.detach()
        a = a + v.detach() * delta
    return (-0.5 * a * a - cf.LOG_SQRT_2PI).sum(-1) + log_det


def score(spec, flat_e, flat_a, s, a, probe, n_steps):
    """flowril.py:183-191: both roles + raw reward."""
    le = log_prob(spec, flat_e, s, a, "expert", probe, n_steps)
    la = log_prob(spec, flat_a, s, a, "agent", probe, n_steps)
    return le, la, le - la


def fm_loss(spec, flat, s, a0, role, t, noise):
    """pretrain.py:179-190 with t / noise injected."""
    views = _views(spec, flat)
    a_t = (1 - t) * a0 + t * noise
    v = velocity(spec, views, a_t, s, t, 1.0 if role == "expert" else -1.0)
    w = (1 - t).pow(1.5)
    return (w * ((v - (noise - a0)) ** 2).sum(1, keepdim=True)).mean()


def train_step(spec, flat, opt, batches, max_norm=1.0):
    """flowril.py:311-330 (no regularisers): losses, backward, clip_grad_norm_, Adam.step."""
    opt.zero_grad()
    losses = [fm_loss(spec, flat, *b) for b in batches]
    sum(losses).backward()
    torch.nn.utils.clip_grad_norm_([flat], max_norm)
    opt.step()
    return [float(l.detach()) for l in losses]
