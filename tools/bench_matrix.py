"""Matrix-path measurement (BASELINE.json configs[3] shape, scaled by arguments): site-mode
divergence matrix of every sample pair = genotype decode + int8 tensor-core contraction.
  python tools/bench_matrix.py [--samples N] [--sites S] [--generations G] [--legacy]
Prints one JSON line: sample*sites/s for the decode, the contraction and the whole call."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tskit_b200.sim import add_mutations, wright_fisher
from tskit_b200.lowlevel import LLTreeSequence

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=20000)
ap.add_argument("--sites", type=int, default=1000000)
ap.add_argument("--generations", type=int, default=2000)
ap.add_argument("--length", type=float, default=1e8)
ap.add_argument("--path", default="", choices=["", "cpasync", "onehot", "legacy"])
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
if args.path:
    os.environ["TSKB_MATRIX"] = args.path
t0 = time.time()
t = wright_fisher(args.samples, args.generations, args.length, ncross=1, seed=11)
add_mutations(t, int(args.sites * 1.02), seed=5)
gen_s = time.time() - t0
n, S = t.num_samples, t.num_sites
ll = LLTreeSequence(t)
L = t.sequence_length
best = None
for _ in range(args.reps):
    t0 = time.perf_counter()
    d = ll.divergence_matrix([0, L], mode="site", span_normalise=False)
    dt = time.perf_counter() - t0
    ph = ll.matrix_phase_ms()
    if best is None or dt < best[0]:
        best = (dt, ph)
dt, ph = best
# property checks that need no oracle: symmetric, zero diagonal, integer counts, and the row sums
# of the biallelic case: sum_j D[i][j] = sum over sites of (carriers of the other allele)
D = d[0]
assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0) and np.array_equal(D, np.round(D))
print(json.dumps({
    "workload": f"site divergence matrix, n={n}, sites={S}, edges={t.num_edges}",
    "impl": {"": "tcgen05 kind::i8, biallelic G G^T, 256x256 tiles, TMA operand delivery, warp-specialised",
             "cpasync": "tcgen05 kind::i8, biallelic G G^T, 256x256 tiles, cp.async operand delivery",
             "onehot": "tcgen05 kind::i8, one-hot per allele",
             "legacy": "mma.sync one-hot"}[args.path],
    "generate_s": round(gen_s, 1), "call_s": dt, "phase_ms": ph,
    "sample_sites_per_s_call": n * S / dt,
    "sample_sites_per_s_decode": n * S / (ph["decode"] / 1e3) if ph["decode"] else None,
    "contraction_TOPS": (2.0 * n * n / 2 * S * ph["alleles"]) / (ph["gemm"] / 1e3) / 1e12 if ph["gemm"] else None,
    "mean_pairwise_diff": float(D.sum() / (n * (n - 1))),
}))
