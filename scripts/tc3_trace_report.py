"""Timeline of one steady-state Euler step from a FRL_TC_TRACE dump of the CTA-pair integrator.
usage: python scripts/tc3_trace_report.py trace.txt [step]"""
import sys
from collections import defaultdict

rows = defaultdict(list)
for line in open(sys.argv[1]):
    r, tag, clk = line.split()
    rows[int(r)].append((int(tag, 16), int(clk)))
step = int(sys.argv[2]) if len(sys.argv) > 2 else 10
names = {0: "epi(leader)", 1: "issuer slot0", 2: "issuer slot1", 3: "producer(leader)"}
# epilogue: marks per step = depth*2 slots*7 + 2 slots*3; find step boundaries by tag 0x100 (l=0, slot=0)
epi = rows[0]
starts = [i for i, (t, _) in enumerate(epi) if t == 0x100]
i0, i1 = starts[step], starts[step + 1]
t0 = epi[i0][1]
print(f"step {step}: {epi[i1][1] - t0} cycles (leader epilogue warp 0)")
for r in (0, 3, 1, 2):
    ev = [(t, c - t0) for t, c in rows[r] if t0 <= c < epi[i1][1]]
    print(f"--- {names[r]}")
    prev = None
    for t, c in ev:
        print(f"   {t:#06x} @{c:7d}" + (f"  (+{c - prev})" if prev is not None else ""))
        prev = c
