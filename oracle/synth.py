"""Deterministic synthetic inputs shared by the golden generator, the tests and bench.py.

TEST INFRASTRUCTURE (see oracle/closed_form.py header).  Everything is drawn from numpy's PCG64, whose stream is
stable across numpy versions, so goldens store seeds instead of megabytes of inputs/weights.

Synthetic data follows SURVEY.md section 8(d): s ~ N(0,1) (VecNormalize'd observations), a ~ clamp(N(0,1), -1, 1),
probe = Rademacher +-1, t = U[0,1) * (1 - 2 eps) + eps, noise ~ N(0,1).
"""
from __future__ import annotations

import numpy as np

# (s_dim, a_dim) per task: reference dims from deps/baselines/get_policy.py:106-127
TASK_DIMS = {
    "sine": (1, 1),
    "maze": (6, 2),
    "hopper": (11, 3),
    "halfcheetah": (17, 6),
    "walker": (17, 6),
    "ant": (132, 8),
    "ant111": (111, 8),
    "push": (16, 3),
    "pick": (16, 4),
    "hand": (68, 20),
}


def rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def states_actions(seed: int, B: int, s_dim: int, a_dim: int):
    g = rng(seed)
    s = g.standard_normal((B, s_dim)).astype(np.float32)
    a = np.clip(g.standard_normal((B, a_dim)), -1.0, 1.0).astype(np.float32)
    return s, a


def rademacher(seed: int, shape) -> np.ndarray:
    return (rng(seed).integers(0, 2, size=shape).astype(np.float32) * 2.0 - 1.0).astype(np.float32)


def cfm_draws(seed: int, B: int, a_dim: int, eps: float = 1e-3, return_u: bool = False):
    """(t [B,1], noise [B,a]) as the reference draws them (pretrain.py:182-183): t = U*(1-2eps)+eps in fp32.
    return_u=True also returns the raw uniform draw (what torch.rand returned inside the reference)."""
    g = rng(seed)
    u = g.random((B, 1), dtype=np.float32)
    t = (u * np.float32(1 - 2 * eps) + np.float32(eps)).astype(np.float32)
    noise = g.standard_normal((B, a_dim)).astype(np.float32)
    return (t, noise, u) if return_u else (t, noise)


def expert_actions(seed: int, s: np.ndarray, a_dim: int) -> np.ndarray:
    """Synthetic 'expert': a = tanh(s[:, :a_dim]) + 0.1 N(0,1); wraps columns if s_dim < a_dim."""
    g = rng(seed)
    cols = np.arange(a_dim) % s.shape[1]
    return (np.tanh(s[:, cols]) + 0.1 * g.standard_normal((s.shape[0], a_dim))).astype(np.float32)
