#!/usr/bin/env python
"""Summarise gpurun_out/<tag>/ into profiles/<tag>_*: the ncu launch list (per-kernel totals and
shares), selected raw metrics of each `--set full` capture, and the bench lines."""
import collections
import csv
import glob
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]


EXTRA = ["sm__pipe_tensor_cycles_active_realtime.avg.pct", "imma_cycles_active", "lts__throughput.avg.pct",
         "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
         "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct"]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("unnamed>::", "")
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none ({path})\n")
        f.write("# cold-cache, serialised launches: compare shares, not absolutes\n")
        f.write(f"{'kernel':50s} {'launches':>8s} {'total_us':>12s} {'mean_us':>10s} {'share':>7s}\n")
        for k, a in agg.items():
            f.write(f"{k:50s} {a[0]:8d} {a[1]:12.1f} {a[1] / a[0]:10.2f} {a[1] / tot:7.3f}\n")
        f.write(f"{'TOTAL':50s} {sum(a[0] for a in agg.values()):8d} {tot:12.1f}\n")


TRAFFIC = {}


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ({path})\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            f.write("\n== " + d.get("Kernel Name", "?")[:150] + "\n")
            try:
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                tot = sum(float(d[k].replace(",", "")) * scale[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                nm = d["Kernel Name"].split("(")[0].replace("void ", "").replace("unnamed>::", "").replace("tskb::<", "").strip()
                TRAFFIC.setdefault(nm, tot)
            except (KeyError, ValueError):
                pass
            for k in KEYS:
                if k in d:
                    f.write(f"{k:70s} {d[k]:>18s} {u[k]}\n")
            for k in d:  # tensor-pipe and L2 figures (names carry a section prefix in newer ncu)
                if any(x in k for x in EXTRA) and d[k] not in ("", "0"):
                    f.write(f"{k:100s} {d[k]:>18s} {u[k]}\n")
        det = subprocess.run(["ncu", "-i", path, "--page", "details"], capture_output=True, text=True).stdout
        keep = ("Duration", "DRAM Throughput", "L2 Hit", "L1/TEX Hit", "Issued Ipc", "Eligible",
                "Warp Cycles Per Issued", "being stalled", "stall type", "Achieved Occ",
                "Theoretical Occ", "Registers Per", "Local", "Mem Busy", "Max Bandwidth")
        f.write("\n== details excerpt\n")
        for ln in det.splitlines():
            if any(k in ln for k in keep):
                f.write(ln.rstrip() + "\n")


if __name__ == "__main__":
    tag = sys.argv[1]
    src = os.path.join("gpurun_out", tag)
    os.makedirs("profiles", exist_ok=True)
    if os.path.exists(os.path.join(src, "launches.csv")):
        launches(os.path.join(src, "launches.csv"), f"profiles/{tag}_launches.txt")
    for rep in glob.glob(os.path.join(src, "*.ncu-rep")):
        full(rep, f"profiles/{tag}_{os.path.basename(rep)[:-8]}.txt")
    if TRAFFIC:
        import json
        p = "profiles/traffic.json"
        cur = json.load(open(p)) if os.path.exists(p) else {}
        cur.update(TRAFFIC)
        cur["_source"] = tag
        json.dump(cur, open(p, "w"), indent=1, sort_keys=True)
    for nm in ("bench.json", "bench_ref.json", "pytest_gpu.log", "smi.txt"):
        p = os.path.join(src, nm)
        if os.path.exists(p):
            open(f"profiles/{tag}_{nm}", "w").write(open(p).read())
