"""Multi-GPU plumbing of the hot path (one process per GPU, torch.distributed over NCCL/NVLink).

SURVEY.md section 8(e):
  * scoring shards by row -- contiguous shards, replicated weights (<= 0.95 MB), NO collective on the data path
    (an optional all_gather returns the [B] outputs to every rank);
  * CFM training is data parallel -- every rank runs the fused gradient kernels on its shard with the GLOBAL batch
    as mean divisor, ONE all-reduce(SUM) of the flat fp32 gradient (< 1 MB: latency-bound on NVLink 5), then the
    same deterministic clip+Adam everywhere, so the weights stay bit-identical without a broadcast.
The functions here are device-agnostic (the gloo tests run them on CPU tensors).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_rows(n_rows: int, rank: int, world: int):
    """Contiguous [start, end) row shard of rank; shards differ by at most one row and cover [0, n_rows)."""
    base, rem = divmod(n_rows, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_flat(grads, losses, group=None):
    """Sum the flat gradient buffers (and the per-term loss scalars) over the data-parallel group, in place."""
    for g in grads:
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
    if losses:
        packed = torch.stack([l.reshape(()) for l in losses])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        for l, v in zip(losses, packed):
            l.copy_(v.reshape(l.shape))


def gather_rows(local: torch.Tensor, n_rows: int, group=None):
    """all_gather of row-sharded outputs ([n_local, ...] per rank, shards from `shard_rows`) -> [n_rows, ...]."""
    world = dist.get_world_size(group)
    sizes = [shard_rows(n_rows, r, world) for r in range(world)]
    max_n = max(e - s for s, e in sizes)
    pad = torch.zeros((max_n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:e - s] for o, (s, e) in zip(out, sizes)], dim=0)


def broadcast_params(nets, src: int = 0, group=None):
    """Make every rank start from rank `src`'s weights (flat buffers of the `_FlatVNet`s)."""
    for net in nets:
        dist.broadcast(net.flat_params(), src=src, group=group)
        net.mark_dirty()


def score_sharded(model, s, a, probe=None, n_steps=32, reward_type="raw", precision=None, gather=False, group=None):
    """Score this rank's shard of a batch that every rank holds in full (or pass already-sharded tensors with
    gather=False and ignore the shard bounds).  Returns (logp_E, logp_A, reward) for the shard, or for all rows when
    gather=True."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = a.shape[0]
    lo, hi = shard_rows(n, rank, world)
    pr = None if probe is None else (probe[lo:hi] if probe.dim() == 2 else probe[:, lo:hi])
    outs = model.score(s[lo:hi], a[lo:hi], n_steps=n_steps, reward_type=reward_type, probe=pr, precision=precision)
    if gather and world > 1:
        return tuple(gather_rows(o, n, group) for o in outs)
    return outs
