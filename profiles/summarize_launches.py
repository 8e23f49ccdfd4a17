#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv [first_kernel_regex]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = r["Kernel Name"]
        grid, block = r.get("Grid Size", ""), r.get("Block Size", "")
        rows.append((int(r["ID"]), name, grid, block, ns))
    agg = OrderedDict()
    for _, name, grid, block, ns in rows:
        short = re.sub(r"\(.*", "", name)
        k = short
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print(f"{len(rows)} launches, total {tot / 1e6:.3f} ms (cold-cache, serialised)")
    for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ns / 1e6:10.3f} ms  {100 * ns / tot:5.1f}%  n={c:5d}  avg {ns / c / 1e3:9.2f} us  {k}")


if __name__ == "__main__":
    main()
