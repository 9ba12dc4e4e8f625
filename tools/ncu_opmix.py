#!/usr/bin/env python
"""Instruction mix and top stall sites of one kernel from an ncu report (source page, SASS view).

    python tools/ncu_opmix.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
si, ii, ti, wi = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
mix, stall = defaultdict(int), []
tot = 0
for r in rows[hi + 1:]:
    if len(r) <= wi:
        continue
    op = r[si].split()
    if not op:
        continue
    name = op[1] if op[0].startswith("@") else op[0]
    name = name.split(".")[0]
    n = int(r[ii] or 0)
    mix[name] += n
    tot += n
    stall.append((int(r[wi] or 0), r[si].strip(), n))
print("total warp instructions executed: %d" % tot)
for k, v in sorted(mix.items(), key=lambda x: -x[1])[:top]:
    print("  %-10s %14d  %5.1f%%" % (k, v, 100.0 * v / tot))
print("top stall sites (samples, executed, SASS):")
for s, src, n in sorted(stall, reverse=True)[:top]:
    print("  %7d %12d  %s" % (s, n, src[:110]))
