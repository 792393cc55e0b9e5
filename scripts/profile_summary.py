"""Turn one round's ncu outputs (gpurun_out/) into the tracked summary under profiles/.

    python scripts/profile_summary.py r01m "title text"

Reads gpurun_out/<tag>_bench.json (un-profiled bench line), gpurun_out/<tag>_launches.csv (the
``--metrics gpu__time_duration.sum`` launch list of the timed region) and gpurun_out/<tag>_tc3.ncu-rep (one
``--set full`` capture of the dominant kernel); writes profiles/<tag>_bench_maze_fp16.json,
profiles/<tag>_bench_launches_ncu.md and refreshes profiles/traffic.json (the ``roofline.traffic`` source of bench.py).
"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__cluster_dim_x", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    tag = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else tag
    bench = json.loads(open(os.path.join(OUT, f"{tag}_bench.json")).read().strip().splitlines()[-1])
    shutil.copy(os.path.join(OUT, f"{tag}_bench.json"), os.path.join(PROF, f"{tag}_bench_maze_fp16.json"))
    md = [f"# Round 1 ({tag}) — {title}", ""]
    md.append(f"`python bench.py` (un-profiled, exit 0): {bench['value'] / 1e6:.2f} M scored samples/s, "
              f"{bench['ms_per_step']:.3f} ms per step, roofline.frac {bench['roofline']['frac']:.3f}; e2e "
              f"{bench['e2e']['value'] / 1e6:.2f} M samples/s; CFM training {bench['train']['value'] / 1e6:.0f} M rows/s "
              f"(HBM frac {bench['train']['hbm']['frac']:.2f}); cpu_baseline {bench['cpu_baseline']['value']:.0f} samples/s "
              f"on {bench['cpu_baseline']['cores']} cores (`profiles/{tag}_bench_maze_fp16.json`).")
    md += ["", "## Launch list of the timed region", "",
           "`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python bench.py "
           "--steps 3 --warmup 3 --no-cpu`", "", "| # | kernel | grid | block | duration (us) |", "|---|---|---|---|---|"]
    lines = [ln for ln in open(os.path.join(OUT, f"{tag}_launches.csv")) if ln.startswith('"')]
    share = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO("".join(lines))):
        us = float(r["Metric Value"]) / 1e3
        name = r["Kernel Name"].split("(")[0][:60]
        md.append(f"| {r['ID']} | {name} | {r['Grid Size']} | {r['Block Size']} | {us:.1f} |")
        share[name] = share.get(name, 0.0) + us
    tot = sum(share.values())
    md += ["", "Share of the timed region:", ""]
    md += [f"* `{k}`: {v:.1f} us = {100 * v / tot:.1f} %" for k, v in sorted(share.items(), key=lambda kv: -kv[1])]
    rep = os.path.join(OUT, f"{tag}_tc3.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, v = rows[0], rows[1], rows[2]
    md += ["", "## `ncu --set full --clock-control none --import-source on -k regex:logprob_tc3 -s 4 -c 1` on "
           "`python bench.py --steps 2 --warmup 3 --no-cpu`", "", "| metric | unit | value |", "|---|---|---|"]
    val = {}
    for k in WANT:
        if k in h:
            i = h.index(k)
            md.append(f"| {k} | {u[i]} | {v[i]} |")
            val[k] = (u[i], v[i])
    rd = float(val["dram__bytes_read.sum"][1]) * UNIT[val["dram__bytes_read.sum"][0]]
    wr = float(val["dram__bytes_write.sum"][1]) * UNIT[val["dram__bytes_write.sum"][0]]
    md += ["", f"DRAM traffic per launch = {rd + wr:.0f} B (algorithmic input bytes + weight images: "
           f"{bench['roofline']['algorithmic_bytes_per_launch']} B); the capture is cold-cache and serialised, its duration "
           f"agrees with the live CUDA-event time of bench.py ({bench['ms_per_step']:.3f} ms per step incl. the reward kernel)."]
    name = f"{tag}_bench_launches_ncu.md"
    open(os.path.join(PROF, name), "w").write("\n".join(md) + "\n")
    tj = os.path.join(PROF, "traffic.json")
    t = json.load(open(tj))
    t["maze:fp16:65536"] = {"bytes_per_launch": int(rd + wr),
                            "source": f"profiles/{name} (ncu --set full: dram__bytes_read.sum {val['dram__bytes_read.sum'][1]} "
                                      f"{val['dram__bytes_read.sum'][0]} + dram__bytes_write.sum {val['dram__bytes_write.sum'][1]} "
                                      f"{val['dram__bytes_write.sum'][0]})"}
    json.dump(t, open(tj, "w"), indent=1)
    print("wrote", name)


if __name__ == "__main__":
    main()
