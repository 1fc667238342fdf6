"""Piece statistics of a cached workload (from the device plan): sizes of the heights, how many
pieces abut the next slot (reductions merged in registers), piece lengths in windows, zero branch
lengths, reference counts."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tskit_b200.lowlevel import LLTreeSequence  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
t, W, _ = bench.load_workload(name)
ll = LLTreeSequence(t)
bp0 = ll.debug_array("q_bp0", np.uint32)
bp1 = ll.debug_array("q_bp1", np.uint32)
bl = ll.debug_array("q_bl", np.float64)
off = ll.debug_array("q_off", np.uint32)
lb = ll.debug_array("level_begin", np.uint32)
pos = ll.debug_array("bp_pos", np.float64)
real = bp1 != 0xFFFFFFFF
n = int(real.sum())
T = len(pos) - 1
print("slots", len(bp0), "real", n, "padding", int((~real).sum()), "heights", len(lb) - 1, "breakpoints", T)
print("height sizes:", np.diff(lb.astype(np.int64)).tolist())
zero = real & (bl == 0)
print("zero branch length: %.4f" % (zero.sum() / n))
merged = real[:-1] & real[1:] & (bp1[:-1] == bp0[1:])
print("pieces abutting the next slot: %.4f  -> reductions per piece ~ %.3f" % (
    merged.sum() / n, (2 * n - merged.sum()) / n))
x0 = pos[bp0[real]]
x1 = pos[np.minimum(bp1[real], T)]
wid = t.sequence_length / W
nw = np.floor(x1 / wid) - np.floor(x0 / wid)
print("windows crossed per piece: mean %.2f, zero %.3f, percentiles 50/90/99:" % (nw.mean(), (nw == 0).mean()),
      np.percentile(nw, [50, 90, 99]).tolist())
ln = bp1[real].astype(np.int64) - bp0[real]
print("piece length in breakpoints: mean %.0f, percentiles 10/50/90/99:" % ln.mean(),
      np.percentile(ln, [10, 50, 90, 99]).tolist())
cnt = np.diff(off.astype(np.int64))
print("refs per real piece: mean %.3f  >3: %.4f  >4: %.4f" % (cnt[real[:len(cnt)]].mean(),
      (cnt[real[:len(cnt)]] > 3).mean(), (cnt[real[:len(cnt)]] > 4).mean()))
# distinct 32-byte sectors of the delta array touched per warp of 32 consecutive slots (start side)
s = (bp0[: len(bp0) // 32 * 32].reshape(-1, 32) // 4)
d = np.array([len(np.unique(r)) for r in s[:: max(1, len(s) // 20000)]])
print("distinct delta sectors per 32 slots (start breakpoints): mean %.1f" % d.mean())
