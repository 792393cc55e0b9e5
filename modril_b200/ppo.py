"""Returns / GAE and advantage normalisation on the device (scope row f4, first part).

Mirrors ``RolloutStorage.compute_returns`` / ``compute_advantages``
(``deps/rl-toolkit/rlf/storage/rollout_storage.py:178-239``) as two functions over the storage's tensors: the T-step python
loop of small torch ops becomes one ``frl_compute_returns`` launch (bit-identical float32 scan), the advantage
normalisation one ``frl_normalise_advantages`` call.  CUDA tensors only; no PyTorch arithmetic, no CPU fallback.
"""
from __future__ import annotations

import torch

from . import _native
from .flowril.networks import _current_stream, _ptr


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
