#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: executed instructions by opcode and the hottest SASS lines.
usage: ncu_sass_summary.py src.csv [units]   (units = divisor, e.g. warp-contacts per launch)"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
h = rows[1]
si, ei, sm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
ops, samples, tot, tots = collections.Counter(), collections.Counter(), 0, 0
lines = []
for r in rows[2:]:
    try:
        e, s = int(float(r[ei])), int(float(r[sm]))
    except Exception:
        continue
    t = r[si].split()
    if not t:
        continue
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += e; samples[op] += s; tot += e; tots += s
    lines.append((s, e, r[0], r[si]))
print("total warp instructions %d (%.1f per unit), samples %d" % (tot, tot / units, tots))
for op, c in ops.most_common(40):
    print("%-10s %12d  %6.1f per unit  %5.1f%% of stall samples" % (op, c, c / units, 100.0 * samples[op] / max(tots, 1)))
print("--- hottest lines by stall samples")
for s, e, a, src in sorted(lines, reverse=True)[:40]:
    print("%6d %10d %s %s" % (s, e, a[-5:], src[:100]))
