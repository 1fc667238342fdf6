"""Per-level timeline of the propagation kernel (TSKB_TRACE=1) on a cached workload."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TSKB_TRACE"] = "1"
import bench
from tskit_b200.lowlevel import LLTreeSequence
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
t, W, _ = bench.load_workload(name)
ll = LLTreeSequence(t)
s = t.samples
windows = np.linspace(0, t.sequence_length, W + 1)
sizes = np.array([len(s)], dtype=np.uint64)
for _ in range(3):
    ll.diversity(sizes, s, windows=windows, mode="branch")
tr = ll.debug_array("trace", np.uint64).reshape(-1, 6).astype(np.int64)
dep = ll.debug_array("tile_dep", np.uint32)
t0 = tr[:, 0].min()
tr = tr - t0
print("tiles", len(tr), "kernel span us", (tr[:, 5].max()) / 1e3)
lv = np.unique(dep)
print("level first_tile ntiles | start(min) waitbeg(med) released(min,max) gathered(max) lookback(max) done(max) | level_time")
prev_done = 0
rows = []
for i, d in enumerate(lv):
    m = dep == d
    x = tr[m]
    rel = x[:, 2]
    rows.append((i, d, m.sum(), x[:, 0].min(), np.median(x[:, 1]), rel.min(), rel.max(), x[:, 3].max(), x[:, 4].max(), x[:, 5].max()))
for r in rows[:12] + rows[40:46] + rows[-6:]:
    i, d, n, st, wb, r0, r1, g, lb, dn = r
    print(f"{i:3d} {d:6d} {n:5d} | {st/1e3:8.1f} {wb/1e3:8.1f}  rel {r0/1e3:8.1f} {r1/1e3:8.1f}  gath {g/1e3:8.1f}  lb {lb/1e3:8.1f}  done {dn/1e3:8.1f} | {(dn - prev_done)/1e3 if i else dn/1e3:6.1f}")
    prev_done = dn
dn = np.array([r[9] for r in rows]); r0 = np.array([r[5] for r in rows]); r1 = np.array([r[6] for r in rows]); g = np.array([r[7] for r in rows]); lb = np.array([r[8] for r in rows])
print("mean per level (us): prev_done->first_release", np.mean(r0[1:] - dn[:-1]) / 1e3, " release spread", np.mean(r1 - r0) / 1e3,
      " last_release->gathered", np.mean(g - r1) / 1e3, " gathered->lookback", np.mean(lb - g) / 1e3, " lookback->done", np.mean(dn - lb) / 1e3,
      " level", np.mean(np.diff(dn)) / 1e3)
mid = (dep > lv[5]) & (dep < lv[-10])
x = tr[mid]
lvl_done_prev = {}
# release time of each tile's level = min release in the level
relmin = {d: tr[dep == d][:, 2].min() for d in lv}
r0t = np.array([relmin[d] for d in dep[mid]])
for name, a in (("wait_end - level_release", x[:, 2] - r0t), ("gathered - wait_end", x[:, 3] - x[:, 2]),
                ("lookback - gathered", x[:, 4] - x[:, 3]), ("done - lookback", x[:, 5] - x[:, 4]),
                ("done - level_release", x[:, 5] - r0t)):
    print(f"{name:28s} p50 {np.percentile(a, 50)/1e3:6.2f} p90 {np.percentile(a, 90)/1e3:6.2f} p99 {np.percentile(a, 99)/1e3:6.2f} max {a.max()/1e3:6.2f} us")
