"""First-contact GPU script: plan arrays vs the numpy model, then statistics vs the reference."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tskit_b200.tables import Tables
from tskit_b200.lowlevel import LLTreeSequence
from oracle import ref
from tests import plan_model

def main():
    for path in sys.argv[1:]:
        t = Tables.load(path)
        print(path, "N", t.num_nodes, "E", t.num_edges, "S", t.num_sites, flush=True)
        t0 = time.time(); ll = LLTreeSequence(t); print("stage s", time.time() - t0, ll.engine_stats(), flush=True)
        if t.num_edges < 60000:
            print("plan diff:", plan_model.compare(ll, t), flush=True)
        r = ref.RefTreeSequence(t)
        samples = t.samples
        n = len(samples)
        L = t.sequence_length
        for W in (1, 10):
            w = np.linspace(0, L, W + 1)
            for mode in ("branch", "site"):
                sets = [samples]
                sizes = np.array([n], dtype=np.uint64)
                a = ll.diversity(sizes, samples, windows=w, mode=mode)
                b = r.one_way("diversity", sets, windows=w, mode=mode)
                err = np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
                print(mode, "W", W, "diversity relerr", err, ll.engine_stats()["last_kernel_ms"], flush=True)
                sets = [samples[: n // 2], samples[n // 2:]]
                sizes = np.array([len(s) for s in sets], dtype=np.uint64)
                idx = np.array([[0, 1], [0, 0]], dtype=np.int32)
                a = ll.divergence(sizes, samples, idx, windows=w, mode=mode)
                b = r.k_way("divergence", sets, idx, windows=w, mode=mode)
                err = np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
                print(mode, "W", W, "divergence relerr", err, flush=True)
        par, cnt = ll.trees_at([0.0, L / 3, L - 1])
        rp, rc = r.trees_at([0.0, L / 3, L - 1])
        print("trees_at parent", np.array_equal(par, rp), "count", np.array_equal(cnt, rc), flush=True)
        t0 = time.time()
        for _ in range(5):
            ll.diversity(np.array([n], dtype=np.uint64), samples, windows=np.linspace(0, L, 101), mode="branch")
        print("5 calls s", time.time() - t0, ll.engine_stats(), flush=True)

main()
