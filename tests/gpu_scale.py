"""Scale probe: stage a big ARG, time each phase of a few statistics, compare with the reference."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tskit_b200.tables import Tables
from tskit_b200.lowlevel import LLTreeSequence
from tskit_b200.sim import add_mutations
from oracle import ref

path = sys.argv[1]
do_ref = len(sys.argv) > 2 and sys.argv[2] == "ref"
t0 = time.time(); t = Tables.load(path); print("load s", time.time() - t0, flush=True)
if t.num_sites == 0:
    add_mutations(t, 1000000, seed=1)
print("N", t.num_nodes, "E", t.num_edges, "S", t.num_sites, flush=True)
t0 = time.time(); ll = LLTreeSequence(t); print("stage s", time.time() - t0, ll.engine_stats(), flush=True)
s = t.samples; n = len(s); L = t.sequence_length
W = np.linspace(0, L, 1001)
sizes1 = np.array([n], dtype=np.uint64)
sets2 = [s[: n // 2], s[n // 2:]]
sizes2 = np.array([len(x) for x in sets2], dtype=np.uint64)
idx = np.array([[0, 1]], dtype=np.int32)
for rep in range(3):
    t0 = time.time(); a = ll.diversity(sizes1, s, windows=W, mode="branch"); dt = time.time() - t0
    print("branch diversity wall ms", dt * 1e3, ll.engine_stats()["last_kernel_ms"][:6], ll.engine_stats()["last_call_ms"], flush=True)
for rep in range(2):
    t0 = time.time(); b = ll.divergence(sizes2, s, idx, windows=W, mode="branch"); dt = time.time() - t0
    print("branch divergence wall ms", dt * 1e3, ll.engine_stats()["last_kernel_ms"][:6], flush=True)
for rep in range(2):
    t0 = time.time(); c = ll.diversity(sizes1, s, windows=W, mode="site"); dt = time.time() - t0
    print("site diversity wall ms", dt * 1e3, ll.engine_stats()["last_kernel_ms"][:6], flush=True)
sets8 = np.array_split(s, 8); sizes8 = np.array([len(x) for x in sets8], dtype=np.uint64)
pairs = np.array([(i, j) for i in range(8) for j in range(i + 1, 8)], dtype=np.int32)
for rep in range(2):
    t0 = time.time(); d = ll.f2(sizes8, s, pairs, windows=W, mode="branch"); dt = time.time() - t0
    print("branch f2 8 sets 28 pairs wall ms", dt * 1e3, ll.engine_stats()["last_kernel_ms"][:6], flush=True)
if do_ref:
    t0 = time.time(); r = ref.RefTreeSequence(t); print("ref init s", time.time() - t0, flush=True)
    t0 = time.time(); ra = r.one_way("diversity", [s], windows=W, mode="branch"); dt = time.time() - t0
    print("REF branch diversity s", dt, "relerr", np.max(np.abs(a - ra) / np.abs(ra)), flush=True)
    t0 = time.time(); rb = r.k_way("divergence", sets2, idx, windows=W, mode="branch"); dt = time.time() - t0
    print("REF branch divergence s", dt, "relerr", np.max(np.abs(b - rb) / np.abs(rb)), flush=True)
    t0 = time.time(); rc = r.one_way("diversity", [s], windows=W, mode="site"); dt = time.time() - t0
    print("REF site diversity s", dt, "relerr", np.max(np.abs(c - rc) / np.maximum(np.abs(rc), 1e-300)), flush=True)
    t0 = time.time(); rd = r.k_way("f2", sets8, pairs, windows=W, mode="branch"); dt = time.time() - t0
    print("REF f2 s", dt, "relerr", np.max(np.abs(d - rd) / np.maximum(np.abs(rd), 1e-300)), flush=True)
