"""Multi-GPU plumbing of the hot path (one process per GPU, torch.distributed over NCCL/NVLink).

SURVEY.md section 8(e):
  * scoring shards by row -- contiguous shards, replicated weights (<= 0.95 MB), NO collective on the data path
    (an optional all_gather returns the [B] outputs to every rank);
  * CFM training is data parallel -- every rank runs the fused gradient kernels on its shard with the GLOBAL batch
    as mean divisor, ONE all-reduce(SUM) of the flat fp32 gradient (< 1 MB: latency-bound on NVLink 5), then the
    same deterministic clip+Adam everywhere, so the weights stay bit-identical without a broadcast.
The functions here are device-agnostic (the gloo tests run them on CPU tensors), except `PeerAllReduce`: the device-side
one-shot all-reduce over NVLink peer memory (csrc/comm.cu) that replaces the NCCL call of the training step, so that a
whole data-parallel update is ONE CUDA graph.
"""
from __future__ import annotations

import ctypes as C

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


class PeerAllReduce:
    """Sum-all-reduce of small fp32 device vectors (the flat gradient, < 1 MB) through peer-mapped exchange buffers:
    two ordinary kernels on the current stream (``frl_comm_allreduce``), graph-capturable, bit-identical result on every
    rank.  ``torch.distributed`` is used ONCE, at construction, to pass the 64-byte memory handles around (host side,
    any backend).  Pass an instance as ``world[0]`` of ``train_step`` in place of the process group.

    Single-process use (tests, one process driving several GPUs): ``PeerAllReduce.local_group(devices, max_floats)``.
    """

    def __init__(self, max_floats: int, device=None, group=None, _connect=True):
        from . import _native

        self._native = _native
        self.group = group
        if _connect:
            self.world = dist.get_world_size(group) if dist.is_initialized() else 1
            self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("PeerAllReduce needs a CUDA device (there is no CPU path)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.max_floats = int(max_floats)
        self._h = None
        if _connect:
            self._connect_group()

    def _create(self):
        L = self._native.lib()
        h = C.c_void_p()
        self._native.check(L.frl_comm_create(self.rank, self.world, self.device.index, self.max_floats, C.byref(h)),
                           "frl_comm_create")
        self._h = h

    def _connect_group(self):
        """Create the local buffer, pass the handles around, map the peers.  A failure on ANY rank (no peer access, IPC
        disabled in a container, ...) raises on EVERY rank, after the same number of collectives, so that callers can fall
        back to NCCL together instead of dead-locking."""
        L = self._native.lib()
        err, mine = None, (C.c_ubyte * 64)()
        try:
            self._create()
            self._native.check(L.frl_comm_handle(self._h, mine), "frl_comm_handle")
        except Exception as e:   # noqa: BLE001 - reported to all ranks below
            err = f"rank {self.rank}: {e}"
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, (err, bytes(mine)), group=self.group)
        else:
            handles[0] = (err, bytes(mine))
        errs = [h[0] for h in handles if h[0]]
        if not errs:
            try:
                blob = (C.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(h[1] for h in handles))
                self._native.check(L.frl_comm_connect(self._h, blob), "frl_comm_connect")
            except Exception as e:   # noqa: BLE001
                err = f"rank {self.rank}: {e}"
            if self.world > 1:
                # every rank has mapped every buffer before the first cell is written -- and learns of a failure elsewhere
                oks = [None] * self.world
                dist.all_gather_object(oks, err, group=self.group)
                errs = [e for e in oks if e]
            elif err:
                errs = [err]
        if errs:
            self.close()
            raise RuntimeError("PeerAllReduce: peer-memory exchange unavailable (" + "; ".join(errs) + ")")

    @classmethod
    def local_group(cls, devices, max_floats: int):
        """One comm per device of THIS process, connected to each other (no torch.distributed needed)."""
        comms = []
        for r, d in enumerate(devices):
            c = cls(max_floats, device=d, _connect=False)
            c.rank, c.world = r, len(devices)
            c._create()
            comms.append(c)
        arr = (C.c_void_p * len(comms))(*[c._h for c in comms])
        for c in comms:
            c._native.check(c._native.lib().frl_comm_connect_local(c._h, arr), "frl_comm_connect_local")
        return comms

    def _call(self, fn, name, tensors):
        ts = [t for t in tensors]
        for t in ts:
            if t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device:
                raise ValueError(f"{name}: contiguous float32 tensors on {self.device} expected")
        ptrs = (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        ns = (C.c_longlong * len(ts))(*[t.numel() for t in ts])
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._native.check(fn(self._h, ptrs, ns, len(ts), C.c_void_p(stream)), name)

    def allreduce(self, tensors):
        """In place: every tensor <- its sum over the ranks (up to 8 tensors per call, all ranks the same sizes)."""
        self._call(self._native.lib().frl_comm_allreduce, "frl_comm_allreduce", tensors)

    def publish(self, tensors):
        self._call(self._native.lib().frl_comm_publish, "frl_comm_publish", tensors)

    def reduce(self, tensors):
        self._call(self._native.lib().frl_comm_reduce, "frl_comm_reduce", tensors)

    def timeouts(self) -> int:
        return int(self._native.lib().frl_comm_timeouts(self._h))

    def close(self):
        if self._h is not None:
            self._native.lib().frl_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def allreduce_peer(comm: PeerAllReduce, grads, losses):
    """`allreduce_flat` through a PeerAllReduce: gradients and loss scalars in ONE publish + reduce pair."""
    comm.allreduce(list(grads) + [l.reshape(-1) for l in losses])


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
