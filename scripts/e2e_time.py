"""A/B of the host-buffer entry point (frl_score_host): zero-copy input reads vs staged copies."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from modril_b200.flowril import score_pair_host  # noqa: E402


def main():
    task, B, n_steps = sys.argv[1] if len(sys.argv) > 1 else "maze", 65536, 20
    dev = torch.device("cuda:0")
    model = bench.make_trained_model(task, dev, steps=20)
    _, _, s, a, probe = bench.make_inputs(task, B, n_steps)
    hs, ha, hp = (torch.from_numpy(x).pin_memory() for x in (s, a, probe))
    houts = tuple(torch.empty(B).pin_memory() for _ in range(3))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        score_pair_host(model.net, model.net, hs, ha, hp, n_steps=n_steps, precision="fp16", out=houts)
    ts = []
    for _ in range(20):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        score_pair_host(model.net, model.net, hs, ha, hp, n_steps=n_steps, precision="fp16", out=houts)
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e3
    print(f"zero_copy={os.environ.get('FRL_HOST_ZEROCOPY', '1')} {task}: median {np.median(ts):.3f} ms  min {ts.min():.3f}  "
          f"-> {B / np.median(ts) * 1e-3:.2f} M samples/s  checksum {float(houts[2].double().sum()):.6f}")


if __name__ == "__main__":
    main()
