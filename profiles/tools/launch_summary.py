#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python launch_summary.py launches.csv"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("fkb::<unnamed>::", "")[:48]
    v = float(row["Metric Value"].replace(",", ""))
    agg.setdefault(name, []).append(v / (1e3 if row["Metric Unit"] in ("ns", "nsecond") else 1.0))
print(f"{'kernel':50s} {'launches':>8s} {'mean us':>10s} {'last us':>10s}")
for k, v in agg.items():
    print(f"{k:50s} {len(v):8d} {sum(v) / len(v):10.1f} {v[-1]:10.1f}")
