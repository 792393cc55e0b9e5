"""Run bench.py on the other BASELINE configurations (one GPU) and write the table under profiles/.

    python scripts/other_configs.py r01n            # on the GPU box; writes gpurun_out/<tag>_cfg_<w>.json + the table
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rows = []
for w in ("hopper", "halfcheetah", "ant", "ant111", "hand", "pick", "push"):
    out = os.path.join(ROOT, "gpurun_out", f"{tag}_cfg_{w}.json")
    with open(out, "w") as f:
        rc = subprocess.call([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", w, "--steps", "5", "--warmup", "3",
                              "--no-cpu"], stdout=f, stderr=subprocess.DEVNULL)
    if rc != 0:
        rows.append(f"| {w} | failed (rc {rc}) |")
        continue
    d = json.loads(open(out).read().strip().splitlines()[-1])
    r, t = d["roofline"], d["train"]
    rows.append(f"| {w} | {d['config']['batch_per_gpu']} | {d['config']['n_steps']} | {d['value'] / 1e6:.2f} M | {d['ms_per_step']:.2f} | "
                f"{r['achieved']:.0f} | {r['frac']:.3f} | {d['e2e']['value'] / 1e6:.2f} M | {t['value'] / 1e6:.0f} M | {t['hbm']['frac']:.2f} |")
    print(rows[-1], flush=True)
md = [f"# Round 1 ({tag}) — the other BASELINE configurations on one B200 (`python bench.py --workload W --steps 5 --warmup 3 --no-cpu`)", "",
      "| workload | rows | ODE steps | scored samples/s | ms/step | TFLOP/s (algorithmic) | of 1630 | e2e samples/s | CFM train rows/s | train HBM frac |",
      "|---|---|---|---|---|---|---|---|---|---|"] + rows
open(os.path.join(ROOT, "gpurun_out", f"{tag}_other_configs.md"), "w").write("\n".join(md) + "\n")
