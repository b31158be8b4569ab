#!/usr/bin/env python
"""profiles/<round>/traffic.json from the ncu summaries: DRAM bytes (read + written) of one launch of the dominant kernel, which
bench.py quotes as roofline.traffic.  usage: ncu_traffic.py profiles/r02 C3 k_weight_lean"""
import json
import os
import re
import sys

d, workload, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
txt = open(os.path.join(d, kernel + "_ncu_raw.txt")).read()
unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = 0.0
for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
    m = re.search(re.escape(name) + r" = ([0-9.]+) (\w+)", txt)
    tot += float(m.group(1)) * unit[m.group(2)]
p = os.path.join(d, "traffic.json")
t = json.load(open(p)) if os.path.exists(p) else {}
t["%s:%s" % (workload, kernel)] = {"bytes": tot, "source": kernel + "_ncu_raw.txt (ncu --set full --clock-control none of the bench command)"}
json.dump(t, open(p, "w"), indent=1)
print(p, t)
