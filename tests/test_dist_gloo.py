"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row sharding, the data-parallel gradient math
(local shard gradient with the GLOBAL batch as mean divisor, all-reduce SUM == full-batch gradient) and the gather of
row-sharded outputs.  The per-shard arithmetic is done by the oracle here (the CUDA kernels need a GPU); the
collective plumbing is the code the product path runs (modril_b200/dist.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from modril_b200.dist import allreduce_flat, gather_rows, shard_rows
        from oracle import closed_form as cf
        from oracle import synth

        spec = cf.NetSpec(6, 2, 32, 2, 2)
        flat = cf.make_params(spec, 3, 1.0).astype(np.float64)
        B = 37  # ragged on purpose
        s, a = synth.states_actions(1, B, 6, 2)
        t, noise = synth.cfm_draws(2, B, 2)
        lo, hi = shard_rows(B, rank, world)
        # local shard gradient with the global batch as mean divisor == (B_local / B) * local-mean gradient
        l_loc, g_loc = cf.fm_loss_and_grad(spec, flat, s[lo:hi], a[lo:hi], "expert", t[lo:hi], noise[lo:hi])
        scale = (hi - lo) / B
        grad = torch.from_numpy(g_loc * scale)
        loss = torch.tensor([l_loc * scale], dtype=torch.float64)
        allreduce_flat([grad], [loss])
        l_full, g_full = cf.fm_loss_and_grad(spec, flat, s, a, "expert", t, noise)
        ok_grad = np.allclose(grad.numpy(), g_full, rtol=1e-10, atol=1e-12)
        ok_loss = abs(float(loss) - l_full) < 1e-12
        # gather of row-sharded outputs
        local = torch.arange(lo, hi, dtype=torch.float32).reshape(-1, 1).repeat(1, 3)
        full = gather_rows(local, B)
        ok_gather = torch.equal(full[:, 0], torch.arange(B, dtype=torch.float32)) and full.shape == (B, 3)
        q.put((rank, ok_grad, ok_loss, ok_gather))
    finally:
        dist.destroy_process_group()


def test_shard_rows_partition():
    from modril_b200.dist import shard_rows
    for n in (0, 1, 7, 128, 1000003):
        for w in (1, 2, 3, 8):
            bounds = [shard_rows(n, r, w) for r in range(w)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            assert all(bounds[i][1] == bounds[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in bounds]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(180)
def test_dp_gradient_allreduce_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] and r[3] for r in res), res
