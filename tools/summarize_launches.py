#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (and grid).
python tools/summarize_launches.py launches.csv [--by-grid]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
by_grid = "--by-grid" in sys.argv
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    if by_grid:
        name += " grid=" + r.get("Grid Size", "")
    rows.append((name, v))
tot = sum(v for _, v in rows)
ours = sum(v for n, v in rows if re.search(r"\bk_[a-z]", n))
print(f"{len(rows)} kernel launches, sum {tot:.2f} ms; libsaragan_b200 kernels (k_*): {ours:.2f} ms = {100 * ours / tot:.1f} %")
agg = collections.defaultdict(lambda: [0.0, 0])
for n, v in rows:
    agg[n][0] += v
    agg[n][1] += 1
print("       ms  share launches  kernel")
for n, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:70]:
    print(f"{v:9.3f} {100 * v / tot:5.1f}% {c:8d}  {n[:110]}")
