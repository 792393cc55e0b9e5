"""Timing of the BCF policy kernels: forward, 32-step sampler, one bc update (loss+grad+clip+Adam)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from modril_b200.bcf import FlowPolicy, FusedAdam, VectorNetwork, bc_step  # noqa: E402


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    s_dim, a_dim = 11, 3
    net = VectorNetwork(s_dim, a_dim).cuda()
    pol = FlowPolicy(net)
    opt = FusedAdam([net], lr=1e-3, eps=1e-12)
    for B in (32, 128, 4096, 65536):
        s = torch.randn(B, s_dim, device="cuda")
        a = torch.randn(B, a_dim, device="cuda").clamp(-1, 1)
        x0 = torch.randn(B, a_dim, device="cuda")
        t = torch.rand(B, device="cuda")
        it = 200 if B <= 4096 else 20
        f = timeit(lambda: net(x0, s, t), it)
        sm = timeit(lambda: pol.get_action(s, x0=x0), max(5, it // 10))
        tr = timeit(lambda: bc_step(pol, opt, s, a, x0=x0, t=t), it)
        flops_fwd = 2 * B * (128 + 128 * 256 + s_dim * 256 + a_dim * 256 + 768 * 256 + 256 * a_dim)
        print(f"B={B:6d} forward {f:.3f} ms  sampler(32) {sm:.3f} ms ({B / sm * 1e-3:.2f} M actions/s)  "
              f"bc_step {tr:.3f} ms ({B / tr * 1e-3:.2f} M rows/s, {6 * flops_fwd / tr * 1e-9:.1f} TFLOP/s)")


if __name__ == "__main__":
    main()
