"""Piece statistics of a cached workload (from the device plan): how many pieces lie inside one
window, run lengths of equal windows in processing order, zero branch lengths, reference counts."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tskit_b200.lowlevel import LLTreeSequence
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
t, W, _ = bench.load_workload(name)
ll = LLTreeSequence(t)
x = ll.debug_array("q_x", np.float64)
xe = ll.debug_array("q_xe", np.float64)
bl = ll.debug_array("q_bl", np.float64)
off = ll.debug_array("q_off", np.uint32)
lb = ll.debug_array("level_begin", np.uint32)
real = xe >= 0
n = real.sum()
print("slots", len(x), "real", n, "padding", (~real).sum(), "heights", len(lb) - 1)
print("height sizes (first 12):", np.diff(lb)[:12], "last:", np.diff(lb)[-5:])
wid = t.sequence_length / W
w0 = np.floor(x / wid).astype(np.int64)
w1 = np.minimum(np.floor(xe / wid), W - 1).astype(np.int64)
same = real & (w0 == w1)
zero = real & (bl == 0)
print("same-window %.3f  zero-bl %.3f  same&nonzero %.3f  multi&nonzero %.3f" % (
    same.sum() / n, zero.sum() / n, (same & ~zero).sum() / n, (real & ~same & ~zero).sum() / n))
ln = (xe - x)[real]
print("piece length percentiles (bp):", np.percentile(ln, [10, 50, 90, 99]).round(1), "mean", ln.mean().round(1))
# runs of equal window among consecutive slots (active = real, nonzero, same-window)
act = same & ~zero
key = np.where(act, w0, -1 - np.arange(len(x)))  # inactive slots never match
brk = np.flatnonzero(key[1:] != key[:-1])
runs = np.diff(np.concatenate([[-1], brk, [len(key) - 1]]))
ract = runs[key[np.concatenate([brk, [len(key) - 1]])] >= 0]
print("active pieces", act.sum(), "runs", len(ract), "mean run", ract.mean().round(2),
      "run percentiles", np.percentile(ract, [50, 90, 99]))
# per warp of 128 consecutive slots: distinct windows among active
k128 = key[: len(key) // 128 * 128].reshape(-1, 128)
a128 = act[: len(key) // 128 * 128].reshape(-1, 128)
d = [(len(np.unique(r[m]))) for r, m in zip(k128[::997], a128[::997])]
print("distinct windows per 128 slots (sampled): mean %.1f" % np.mean(d), "active per 128: %.1f" % a128[::997].sum(1).mean())
cnt = np.diff(off.astype(np.int64))
print("refs per real piece: mean %.2f  >3: %.4f" % (cnt[real].mean(), (cnt[real] > 3).mean()))
