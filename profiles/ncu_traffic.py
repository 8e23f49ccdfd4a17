"""Write profiles/roofline_traffic.json from a raw ncu page (`ncu -i x.ncu-rep --page raw --csv > raw.csv`) of the dominant kernel.

    python profiles/ncu_traffic.py profiles/<raw>.csv --kernel-regex 'conv_tc_kernel<256, 3, false, false, 2>' --min-us 60 \
        --key nice_conv2 --precision fp32 --batch 64 --build <git sha>

For every captured launch of the matching kernel whose duration is at least --min-us (the NICE conv2 launches; conv1 shares the template
but is 3x shorter) it takes dram__bytes_read.sum + dram__bytes_write.sum and stores the mean per launch, with the capture's provenance.
bench.py reads the file for `roofline.traffic` (null when the entry does not match the benchmarked precision / batch).
"""
import argparse
import csv
import datetime
import json
import os
import re

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--kernel-regex", required=True)
ap.add_argument("--min-us", type=float, default=0.0)
ap.add_argument("--key", default="nice_conv2")
ap.add_argument("--precision", default="fp32")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--build", default="")
a = ap.parse_args()

with open(a.csv, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.reader(lines))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
need = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"]
for n in need:
    assert n in col, f"column {n} missing from the raw page"
opt = [n for n in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "lts__t_bytes.sum") if n in col]


def scale(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)


pat = re.compile(a.kernel_regex)
sel = []
for r in rows[2:]:
    if len(r) < len(hdr) or not pat.search(r[col["Kernel Name"]]):
        continue
    us = scale(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]])
    if us < a.min_us:
        continue
    rd = scale(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
    wr = scale(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    sel.append((us, rd, wr, {n: r[col[n]] for n in opt}))
assert sel, "no launch matched"
n = len(sel)
entry = {
    "kernel": a.kernel_regex, "launches_averaged": n, "precision": a.precision, "batch": a.batch,
    "gpu_time_us": sum(s[0] for s in sel) / n, "dram_read_bytes": sum(s[1] for s in sel) / n, "dram_write_bytes": sum(s[2] for s in sel) / n,
    "dram_bytes_per_launch": sum(s[1] + s[2] for s in sel) / n, "extra": sel[0][3],
    "source": os.path.basename(a.csv), "build": a.build, "written": datetime.date.today().isoformat(),
    "note": "ncu --set full --clock-control none: cold-cache, serialised launches; per launch like roofline.achieved",
}
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "roofline_traffic.json")
data = {}
if os.path.exists(out):
    data = json.load(open(out))
data[a.key] = entry
json.dump(data, open(out, "w"), indent=1)
print(json.dumps(entry, indent=1))
