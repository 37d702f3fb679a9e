#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): launches, total time and
share per kernel.  python scripts/launch_summary.py profiles/r2_launches_contacts.csv "<command>" """
import csv
import re
import sys

path = sys.argv[1]
command = sys.argv[2] if len(sys.argv) > 2 else ""
lines = [line for line in open(path, encoding="utf-8") if line.startswith('"')]
rows = list(csv.DictReader(lines))
ours, other = {}, {}
for row in rows:
    if row["Metric Name"] != "gpu__time_duration.sum":
        continue
    unit = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}[row["Metric Unit"]]
    ms = float(row["Metric Value"].replace(",", "")) * unit
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
    table = ours if "cvp::" in row["Kernel Name"] or name.startswith("stream_fused") else other
    count, total = table.get(name, (0, 0.0))
    table[name] = (count + 1, total + ms)
print(f"ncu --metrics gpu__time_duration.sum --clock-control none: {command}")
print(f"(launch list: {path.split('/')[-1]}; per-launch times are cold-cache and serialised -- read the SHARES)\n")
total = sum(t for _, t in ours.values())
print(f"{'kernel':90s} {'launches':>8s} {'total ms':>10s} {'share of cvp kernels':>22s}")
for name, (count, ms) in ours.items():
    print(f"{name[:90]:90s} {count:8d} {ms:10.3f} {ms / total:22.4f}")
print()
for name, (count, ms) in other.items():
    print(f"[torch setup] {name[:76]:76s} {count:8d} {ms:10.3f}")
