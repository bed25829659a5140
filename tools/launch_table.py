#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches and mean duration per (kernel, grid).
    python tools/launch_table.py profiles/r2_launches_bench_default.csv > table.md"""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    agg.setdefault((name, r[gi], r[bi]), []).append(float(r[vi].replace(",", "")) / 1e3)
print("| kernel | grid | block | launches | mean us | min us | max us |")
print("|---|---|---|---|---|---|---|")
for (name, grid, block), v in agg.items():
    print(f"| `{name}` | {grid} | {block} | {len(v)} | {sum(v) / len(v):.1f} | {min(v):.1f} | {max(v):.1f} |")
