"""Latency of the device-side gradient all-reduce (csrc/comm.cu) next to NCCL's, for the flat gradient of the bench net.

    python scripts/comm_time.py                                   # one GPU: the two kernels by themselves (world 1)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/comm_time.py

Prints one JSON line on rank 0: microseconds per all-reduce (device time, max over ranks) replayed from a CUDA graph of 50
back-to-back all-reduces (peer) and launched eagerly (NCCL; captured NCCL made the process hang at exit).
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from modril_b200.dist import PeerAllReduce  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 202_000   # maze H = 256 depth 4 net: 0.8 MB of fp32 gradient
    comm = PeerAllReduce(n + 64, device=dev)
    g = torch.randn(n, device=dev) * 1e-3
    l0, l1 = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
    reps = 50

    def timed(fn, iters):
        fn()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t) * 1e3

    for _ in range(3):
        comm.allreduce([g, l0, l1])
    torch.cuda.synchronize(dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            comm.allreduce([g, l0, l1])
    peer_us = timed(graph.replay, 10) / (10 * reps)
    out = {"world": world, "floats": n, "peer_allreduce_us": peer_us, "timeouts": comm.timeouts()}
    if world > 1:
        def nccl():
            for _ in range(reps):
                dist.all_reduce(g)
        out["nccl_allreduce_us"] = timed(nccl, 10) / (10 * reps)
    if rank == 0:
        print(json.dumps(out), flush=True)
    comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
