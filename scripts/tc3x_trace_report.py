"""Timeline of one Euler step from a FRL_TC_TRACE dump of the split-precision CTA-pair integrator (logprob_tc3x.cu).
usage: python scripts/tc3x_trace_report.py trace.txt
roles: 0 epilogue warp 0 of the leader (0x1LL wait d_full, 0x2LL got it, 0x5LL accumulators in registers, 0x3LL operand
written; LL = 2*layer + half; 0x1F0/0x2F0/0x3F0 heads), 1 / 2 first issuer of the main / small-terms stream (0x6BB
readiness observed, 0x3BB weights landed + baton, 0x7BB MMAs issued, 0x4BB committed; BB = block), 3 weight producer (0x2BB slot free).
Stamps are 32-bit clocks of the same SM."""
import sys
from collections import defaultdict

rows = defaultdict(list)
for line in open(sys.argv[1]):
    r, tag, clk = line.split()
    if 0 < int(tag, 16) < 0x1000:
        rows[int(r)].append((int(tag, 16), int(clk)))
t0 = min(c for r_, ev in rows.items() if r_ < 4 for _, c in ev)
t0_peer = min([c for _, c in rows.get(4, [])] or [0])
names = {0: "epilogue warp 0 (leader)", 1: "issuer main stream", 2: "issuer small stream", 3: "producer (leader)",
         4: "epilogue warp 0 (PEER CTA: another SM clock -- compare durations, not absolute stamps)"}
for r in (0, 1, 2, 3, 4):
    print(f"--- {names[r]}")
    prev = None
    for t, c in rows[r]:
        c = (c - (t0_peer if r == 4 else t0)) & 0xFFFFFFFF
        print(f"   {t:#06x} @{c:7d}" + (f"  (+{c - prev})" if prev is not None else ""))
        prev = c
