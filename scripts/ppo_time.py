"""Timing of one PPO.update of the reference's default shape (4096 rows, 4 x 4 minibatches, 64 x 2 nets): eager vs CUDA
graph; under `ncu --metrics gpu__time_duration.sum` the eager pass gives the per-kernel launch list."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from modril_b200.ppo import DistActorCritic, FusedAdam, PPOUpdateGraph, ppo_update

dev = torch.device("cuda:0")
hid = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
pol = DistActorCritic(11, 3, hid, 2).to(dev)
obs = torch.randn(rows, 11, device=dev)
out = pol.get_action(obs)
data = (obs, out["action"], out["action_log_probs"], out["value"], out["value"] + 0.1 * torch.randn_like(out["value"]),
        torch.randn(rows, 1, device=dev))
opt = FusedAdam([pol], lr=1e-3, eps=1e-12, capturable=True)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print(f"eager  {timed(lambda: ppo_update(pol, opt, *data)):.3f} ms per update (16 minibatches)")
g = PPOUpdateGraph(pol, opt, rows)
print(f"graph  {timed(lambda: g(*data)):.3f} ms per update")
torch.cuda.profiler.start()
ppo_update(pol, opt, *data, num_epochs=1, num_mini_batch=4)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
