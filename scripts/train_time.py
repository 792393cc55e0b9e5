"""Time N CFM train steps (both roles, B rows each) in one precision; short enough to run under ncu.
usage: train_time.py TASK B PRECISION STEPS [graph]     (graph: capture one step into a CUDA graph and replay it)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modril_b200.flowril import CoupledResidualFM, FusedAdam, train_step
from oracle import synth

task, B, prec, steps = sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
use_graph = len(sys.argv) > 5 and sys.argv[5] == "graph"
s_dim, a_dim = synth.TASK_DIMS[task]
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = CoupledResidualFM(s_dim, a_dim, 256).to(dev)
opt = FusedAdam([model.net], lr=1e-4, capturable=use_graph)
s, a = (torch.from_numpy(x).to(dev) for x in synth.states_actions(1, B, s_dim, a_dim))
g = torch.Generator().manual_seed(1)
t = (torch.rand(2, B, 1, generator=g) * 0.998 + 1e-3).to(dev)
z = torch.randn(2, B, a_dim, generator=g).to(dev)
def step():
    return train_step([dict(net=model.net, sign=1.0, s=s, a0=a, t=t[0], noise=z[0]),
                       dict(net=model.net, sign=-1.0, s=s, a0=a, t=t[1], noise=z[1])], opt, 1.0, precision=prec)
for _ in range(3):
    step()
torch.cuda.synchronize()
if use_graph:
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        losses = step()
    run = graph.replay
else:
    run = step
run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"{task} B={B} {prec}{' graph' if use_graph else ''}: {ms:.3f} ms/step, {2 * B / ms / 1e3:.2f} M rows/s, "
      f"adam steps {opt.step_count()}, stage time-outs {model.net.train_timeouts()}, losses {[float(x) for x in step()]}", flush=True)
