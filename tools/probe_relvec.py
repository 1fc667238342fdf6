"""Timing of genetic_relatedness_vector (GRM x vector, branch mode) on the cached C2 ARG: the C-ABI call
with host buffers on one B200 (best of 3), and the reference package's own call on the same tables on
one host core, with the largest relative difference between the two results."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import bench
from tskit_b200.lowlevel import LLTreeSequence


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return r, best * 1e3


t, W, _ = bench.load_workload("c2")
ll = LLTreeSequence(t)
L, s = t.sequence_length, t.samples
n = len(s)
rng = np.random.default_rng(1)
out = {"num_samples": n, "num_edges": int(t.num_edges)}
res = {}
QUICK = os.environ.get("PROBE_QUICK") is not None  # one call only (under ncu)
for K in ((1,) if QUICK else (1, 8)):
    Wt = rng.normal(size=(n, K))
    for nw in ((1,) if QUICK else (1, 10)):
        w = np.linspace(0, L, nw + 1)
        r, ms = timed(lambda: ll.genetic_relatedness_vector(Wt, w, mode="branch", centre=True, nodes=s),
                      reps=1 if QUICK else 3)
        out[f"relvec_{K}cols_{nw}windows_ms"] = ms
        res[(K, nw)] = (Wt, w, r)
        st = ll.engine_stats()
        out[f"relvec_{K}cols_{nw}windows_launches"] = int(st["last_launches"])
        out[f"relvec_{K}cols_{nw}windows_device_ms"] = float(st["last_call_ms"])
# the C-ABI call alone (arrays prepared once), to separate the ctypes mirror's share of the wall time
import ctypes as C
from tskit_b200 import _lib
Wt, w, _ = res[(1, 1)]
focal = np.ascontiguousarray(s, dtype=np.int32)
result = np.zeros((1, n, 1))
p = lambda a: a.ctypes.data_as(C.c_void_p)
_, out["relvec_1cols_1windows_cabi_only_ms"] = timed(lambda: _lib.lib().tskb_treeseq_genetic_relatedness_vector(
    ll._h, 1, p(Wt), 1, p(w), n, p(focal), p(result), 2), reps=5)
out["relvec_1cols_1windows_phase_ms"] = [float(x) for x in ll.engine_stats()["last_kernel_ms"]]
try:
    if QUICK:
        raise RuntimeError("quick run")
    from tskit_b200 import dropin
    ts = dropin.from_tables(t)
    for key in ((1, 1), (8, 1)):
        Wt, w, got = res[key]
        t0 = time.perf_counter()
        want = ts.genetic_relatedness_vector(Wt, windows=w, mode="branch", centre=True)
        out[f"reference_{key[0]}cols_{key[1]}windows_ms"] = (time.perf_counter() - t0) * 1e3
        out[f"max_rel_diff_{key[0]}cols"] = float(np.abs(got - want).max() / np.abs(want).max())
except Exception as e:  # the reference package did not travel
    out["reference"] = f"unavailable: {e!r}"
print(json.dumps(out))
