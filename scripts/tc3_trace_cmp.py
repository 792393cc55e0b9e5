"""Side by side: marks of trace roles 0 and 3 (two epilogue warps) for one step."""
import sys
from collections import defaultdict
rows = defaultdict(list)
for line in open(sys.argv[1]):
    r, tag, clk = line.split()
    rows[int(r)].append((int(tag, 16), int(clk)))
step = int(sys.argv[2]) if len(sys.argv) > 2 else 10
epi = rows[0]
starts = [i for i, (t, _) in enumerate(epi) if t == 0x100]
t0, t1 = epi[starts[step]][1], epi[starts[step + 1]][1]
a = {t: c - t0 for t, c in epi if t0 <= c < t1}
b = {}
for t, c in rows[3]:
    if t0 - 500 <= c < t1 + 500 and t not in b:
        b[t] = c - t0
for t in sorted(a, key=lambda k: a[k]):
    print(f"{t:#06x}  warp0 @{a[t]:6d}   other @{b.get(t, -1):6d}   diff {b.get(t, 0) - a[t]:5d}")
