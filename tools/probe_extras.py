"""Timings of the statistics beyond the headline path on the cached C2 ARG (one B200, wall clock of the
C-ABI call with host buffers, best of 3): weighted statistics, node mode, allele frequency spectra."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
w = np.linspace(0, L, W + 1)
w10 = np.linspace(0, L, 11)
rng = np.random.default_rng(1)
Wt = rng.normal(size=(n, 2))
out = {}
for mode in ("branch", "site"):
    _, out[f"trait_covariance_2cols_{mode}_ms"] = timed(lambda: ll.trait_covariance(Wt, w, mode=mode, span_normalise=True))
    _, out[f"trait_correlation_2cols_{mode}_ms"] = timed(lambda: ll.trait_correlation(Wt, w, mode=mode, span_normalise=True))
    _, out[f"relatedness_weighted_3pairs_{mode}_ms"] = timed(lambda: ll.genetic_relatedness_weighted(
        Wt, np.array([[0, 0], [0, 1], [1, 1]], dtype=np.int32), w, mode=mode))
sz = np.array([n], dtype=np.uint64)
for mode in ("branch", "site"):
    r, ms = timed(lambda: ll.allele_frequency_spectrum(sz, s, w10, [0, np.inf], mode=mode, polarised=True))
    out[f"afs_1set_10windows_{mode}_ms"] = ms
    out[f"afs_1set_10windows_{mode}_result_MB"] = r.nbytes / 1e6
sets2 = np.array([300, 400], dtype=np.uint64)
r, ms = timed(lambda: ll.allele_frequency_spectrum(sets2, s[:700], w, [0, np.inf], mode="branch"))
out["joint_afs_300x400_1000windows_branch_ms"] = ms
t0 = time.perf_counter(); r = ll.diversity(sz, s, windows=w10, mode="node"); first = (time.perf_counter() - t0) * 1e3
_, ms = timed(lambda: ll.diversity(sz, s, windows=w10, mode="node"))
out["node_diversity_10windows_ms"] = ms
out["node_diversity_first_call_ms_incl_second_plan"] = first
out["node_result_MB"] = r.nbytes / 1e6
print(json.dumps(out))
