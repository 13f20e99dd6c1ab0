#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total, share.

    python profiles/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches.txt
"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = r[ki].split("(")[0]
    agg.setdefault((name, r[gi]), []).append(v)
total = sum(sum(v) for v in agg.values())
print(f"{'kernel':58s} {'grid':>16s} {'n':>4s} {'avg us':>9s} {'sum us':>10s} {'share':>6s}")
for (name, grid), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{name[:58]:58s} {grid:>16s} {len(v):4d} {sum(v) / len(v) / 1e3:9.1f} {sum(v) / 1e3:10.1f} {100 * sum(v) / total:5.1f}%")
