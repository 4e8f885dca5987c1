#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
skip = tuple(sys.argv[2:])          # kernel-name substrings to leave out (setup kernels)
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("b200::", "")
    if any(s in name for s in skip):
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-52s n=%5d  total=%10.1f us  share=%5.1f%%  avg=%9.1f us" % (k[:52], v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
print("TOTAL %.1f us over %d launches" % (tot, sum(v[0] for v in agg.values())))
