"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share, launches and average duration per kernel.

    python profiles/summarize_launches.py profiles/<launches>.csv [top_n]
"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
mi = hdr.index("Metric Name")
acc = defaultdict(lambda: [0, 0.0])
for row in r:
    if len(row) <= vi or row[mi] != "gpu__time_duration.sum":
        continue
    v = float(row[vi].replace(",", ""))
    u = row[ui]
    us = v / 1000.0 if u in ("nsecond", "ns") else (v if u in ("usecond", "us") else v * 1000.0)
    name = re.sub(r"\(.*$", "", row[ki])
    name = re.sub(r"^void ", "", name).replace("ipk::", "").replace("(int)", "").replace("(bool)", "")
    acc[name][0] += 1
    acc[name][1] += us
tot = sum(v[1] for v in acc.values())
n = sum(v[0] for v in acc.values())
print(f"{n} launches, {tot / 1000.0:.2f} ms of kernel time")
print("| share | kernel | launches | avg us | total ms |\n|---|---|---|---|---|")
for k, (c, t) in sorted(acc.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"| {100 * t / tot:.1f} % | `{k}` | {c} | {t / c:.1f} | {t / 1000:.2f} |")
