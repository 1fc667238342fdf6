#!/usr/bin/env python
"""profiles/<tag>_relvec_launches.txt from gpurun_out/<tag>/relvec_launches.csv (the ncu launch list of one
genetic_relatedness_vector call, written by tools/gpu_round.sh)."""
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", tag, "relvec_launches.csv"))))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hd = rows[h]
ix = {n: hd.index(n) for n in ("ID", "Kernel Name", "Grid Size", "Metric Name", "Metric Unit", "Metric Value")}
SCALE = {"us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
L = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) < len(hd):
        continue
    d = L.setdefault(r[ix["ID"]], {"k": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]]})
    d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * SCALE.get(r[ix["Metric Unit"]], 1.0)


def short(k):
    for n in ("k_init_weights", "k_sweep", "k_relvec_push", "k_relvec_out"):
        if n in k:
            return n
    return k[:30]


out = ["# PROBE_QUICK=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none",
       f"#   -k regex:(k_relvec|k_sweep|k_init_weights) python tools/probe_relvec.py   (gpurun_out/{tag}/relvec_launches.csv)",
       "# one genetic_relatedness_vector call on C2 (n = 1e5, E = 1e7, 49.1 M pieces, 39 heights), 1 weight column, 1 window;",
       "# cold-cache, serialised launches: compare shares, not absolutes.  Push launches are listed tallest height first.",
       f"{'id':>3} {'kernel':16} {'grid':>14} {'us':>10} {'read_MB':>10} {'write_MB':>10}"]
tot = collections.OrderedDict()
first = None
for i, d in L.items():
    k = short(d["k"])
    if k == "k_init_weights":
        if first is not None:
            break  # the launch list of the first call only
        first = i
    t = d.get("gpu__time_duration.sum", 0) / 1e3
    rd, wr = d.get("dram__bytes_read.sum", 0) / 1e6, d.get("dram__bytes_write.sum", 0) / 1e6
    out.append(f"{i:>3} {k:16} {d['grid']:>14} {t:10.2f} {rd:10.2f} {wr:10.2f}")
    a = tot.setdefault(k, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += t; a[2] += rd; a[3] += wr
out += ["", f"{'kernel':16} {'launches':>8} {'total_us':>10} {'read_MB':>10} {'write_MB':>10}"]
out += [f"{k:16} {a[0]:8d} {a[1]:10.2f} {a[2]:10.2f} {a[3]:10.2f}" for k, a in tot.items()]
dst = os.path.join(ROOT, "profiles", f"{tag}_relvec_launches.txt")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[-6:]))
