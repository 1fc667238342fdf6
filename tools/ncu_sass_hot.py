#!/usr/bin/env python
"""Hot SASS of one kernel in an .ncu-rep: instruction-class histogram and top stall lines.
usage: ncu_sass_hot.py report.ncu-rep [kernel-index]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(out.splitlines()):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": [], "hdr": None}
        blocks.append(cur)
    elif cur is not None and row and row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] and len(row) == len(cur["hdr"]):
        cur["rows"].append(row)
b = blocks[which]
h = b["hdr"]
si, ii, st = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
print(b["name"][:120])
tot_i = sum(int(r[ii]) for r in b["rows"])
tot_s = sum(int(r[st]) for r in b["rows"])
ops = collections.Counter()
for r in b["rows"]:
    op = r[si].split()
    op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
    ops[op.split(".")[0]] += int(r[ii])
print("total warp instr", tot_i, "stall samples", tot_s)
print("by opcode:", ", ".join(f"{k} {v / tot_i:.1%}" for k, v in ops.most_common(18)))
print("top stall lines:")
for r in sorted(b["rows"], key=lambda r: -int(r[st]))[:25]:
    print(f"  {int(r[st]) / max(tot_s, 1):6.1%}  exec {int(r[ii]):>10d}  {r[si].strip()[:100]}")
