"""Per-height timeline of the propagation kernel (TSKB_TRACE=1) on a cached workload."""
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
tr = ll.debug_array("trace", np.uint64).reshape(-1, 4).astype(np.int64)
dep = ll.debug_array("tile_dep", np.uint32)
tr = tr - tr[:, 0].min()
print("tiles", len(tr), "kernel span us", tr[:, 3].max() / 1e3, "phases ms", ll.engine_stats()["last_kernel_ms"][:4])
lv = np.unique(dep)
rows = []
for i, d in enumerate(lv):
    x = tr[dep == d]
    rows.append((i, len(x), x[:, 0].min(), x[:, 1].min(), x[:, 1].max(), x[:, 2].max(), x[:, 3].max()))
prev = 0
for r in rows:
    i, n, st, r0, r1, g, dn = r
    if i < 6 or i % 6 == 0 or i > len(rows) - 4:
        print(f"h{i:3d} tiles {n:5d} | first start {st/1e3:8.1f} released {r0/1e3:8.1f}..{r1/1e3:8.1f} gathered {g/1e3:8.1f} done {dn/1e3:8.1f} | +{(dn - prev)/1e3:6.1f}")
    prev = dn
dn = np.array([r[6] for r in rows]); r0 = np.array([r[3] for r in rows]); r1 = np.array([r[4] for r in rows]); g = np.array([r[5] for r in rows])
print("mean per height (us): done->next release %.2f  release spread %.2f  release->gathered %.2f  gathered->done %.2f  height %.2f" % (
    np.mean(r0[1:] - dn[:-1]) / 1e3, np.mean(r1 - r0) / 1e3, np.mean(g - r1) / 1e3, np.mean(dn - g) / 1e3, np.mean(np.diff(dn)) / 1e3))
relmin = {d: tr[dep == d][:, 1].min() for d in lv}
r0t = np.array([relmin[d] for d in dep])
for nm, a in (("wait_end - release", tr[:, 1] - r0t), ("gathered - wait_end", tr[:, 2] - tr[:, 1]), ("done - gathered", tr[:, 3] - tr[:, 2]), ("done - release", tr[:, 3] - r0t)):
    print(f"{nm:22s} p50 {np.percentile(a, 50)/1e3:6.2f} p90 {np.percentile(a, 90)/1e3:6.2f} p99 {np.percentile(a, 99)/1e3:6.2f} max {a.max()/1e3:6.2f} us")
