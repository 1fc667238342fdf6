"""Timings of the BASELINE.json config shapes that are not the bench line (one B200, device timers
and wall clock), on the cached C2 ARG unless stated.  Prints one JSON object."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tskit_b200.lowlevel import LLTreeSequence
from tskit_b200.sim import add_mutations, wright_fisher
from oracle import ref


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return r, best * 1e3


out = {}
# C1: site diversity, 1k samples, 10 Mb, 100 windows (the reference's own CPU-runnable case)
t1 = add_mutations(wright_fisher(1000, 4000, 1e7, ncross=1, seed=3), 20000, seed=4).ensure_derived()
ll1 = LLTreeSequence(t1)
w1 = np.linspace(0, t1.sequence_length, 101)
s1 = t1.samples
sz1 = np.array([len(s1)], dtype=np.uint64)
g, ms = timed(lambda: ll1.diversity(sz1, s1, windows=w1, mode="site"), 5)
c1 = {"edges": int(t1.num_edges), "sites": int(t1.num_sites), "gpu_ms": ms, "phases_ms": ll1.engine_stats()["last_kernel_ms"][:6]}
if ref.available():
    r1 = ref.RefTreeSequence(t1)
    rr, rms = timed(lambda: r1.one_way("diversity", [s1], windows=w1, mode="site"), 3)
    c1["reference_1core_ms"] = rms
    c1["max_rel_err"] = float(np.max(np.abs(g - rr) / np.maximum(np.abs(rr), 1e-300)))
out["c1_site_diversity_1k_10Mb_100w"] = c1

t, W, _ = bench.load_workload("c2")
ll = LLTreeSequence(t)
L = t.sequence_length
s = t.samples
n = len(s)
w = np.linspace(0, L, W + 1)
sz = np.array([n], dtype=np.uint64)
# site mode on the C2 ARG
_, ms = timed(lambda: ll.diversity(sz, s, windows=w, mode="site"))
out["c2_site_diversity"] = {"sites": int(t.num_sites), "gpu_ms": ms, "phases_ms": ll.engine_stats()["last_kernel_ms"][:6],
                            "sample_sites_per_s": n * t.num_sites / (ms / 1e3)}
# C3 shape (scaled to the C2 ARG): 8 sample sets, Fst = 28 divergences + 8 diversities, f2/f3/f4
sets = np.array_split(s, 8)
sizes = np.array([len(x) for x in sets], dtype=np.uint64)
flat = np.concatenate(sets).astype(np.int32)
pairs = np.array([(i, j) for i in range(8) for j in range(i + 1, 8)], dtype=np.int32)
c3 = {}
for mode in ("branch", "site"):
    _, ms = timed(lambda: ll.divergence(sizes, flat, pairs, windows=w, mode=mode))
    c3[f"divergence_28_pairs_{mode}_ms"] = ms
    c3[f"divergence_28_pairs_{mode}_phases_ms"] = ll.engine_stats()["last_kernel_ms"][:6]
    _, ms = timed(lambda: ll.diversity(sizes, flat, windows=w, mode=mode))
    c3[f"diversity_8_sets_{mode}_ms"] = ms
    _, ms = timed(lambda: ll.f4(sizes, flat, np.array([(0, 1, 2, 3), (4, 5, 6, 7), (0, 2, 4, 6)], dtype=np.int32), windows=w, mode=mode))
    c3[f"f4_3_tuples_{mode}_ms"] = ms
out["c3_shape_on_c2_arg_8_sets"] = c3
# C5 shape: sample_count_stat with a custom summary over 10^6 windows
tiny = np.linspace(0, L, 1_000_001)
f = lambda x: x * (n - x) / (n * (n - 1))  # noqa: E731
Wt = np.ones((n, 1))
t0 = time.perf_counter(); r = ll.general_stat(Wt, f, 1, windows=tiny, mode="branch"); first = (time.perf_counter() - t0) * 1e3
_, ms = timed(lambda: ll.diversity(sz, s, windows=tiny, mode="branch", polarised=True))
out["c5_custom_summary_1e6_windows"] = {"general_stat_wall_ms_incl_python_table": first,
                                        "device_call_ms_same_windows": ms,
                                        "phases_ms": ll.engine_stats()["last_kernel_ms"][:6]}
print(json.dumps(out))
