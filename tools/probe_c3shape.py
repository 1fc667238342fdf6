"""Many-column branch statistics on the C2 ARG (the C3 shape: 8 sample sets), window runs against
per-breakpoint deltas (TSKB_COLS_VARIANT=d), device phases and agreement.  Prints one JSON object."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tskit_b200.lowlevel import LLTreeSequence

t, W, _ = bench.load_workload("c2")
ll = LLTreeSequence(t)
s = t.samples
w = np.linspace(0, t.sequence_length, W + 1)
sets = np.array_split(s, 8)
sizes = np.array([len(x) for x in sets], dtype=np.uint64)
flat = np.concatenate(sets).astype(np.int32)
pairs = np.array([(i, j) for i in range(8) for j in range(i + 1, 8)], dtype=np.int32)
triples = np.array([(i, (i + 1) % 8, (i + 2) % 8) for i in range(8)], dtype=np.int32)
quads = np.array([(0, 1, 2, 3), (4, 5, 6, 7), (0, 2, 4, 6)], dtype=np.int32)
calls = {"divergence_28": lambda: ll.divergence(sizes, flat, pairs, windows=w, mode="branch"),
         "f2_28": lambda: ll.f2(sizes, flat, pairs, windows=w, mode="branch"),
         "f3_8": lambda: ll.f3(sizes, flat, triples, windows=w, mode="branch"),
         "diversity_8": lambda: ll.diversity(sizes, flat, windows=w, mode="branch"),
         "f4_3": lambda: ll.f4(sizes, flat, quads, windows=w, mode="branch")}
out = {}
res = {}
for variant in sys.argv[1:] or ["default", "TSKB_COLS_VARIANT=d"]:
    kv = dict(x.split("=") for x in variant.split(",") if "=" in x)
    os.environ.update(kv)
    o = {}
    for name, fn in calls.items():
        fn()
        best, ph = None, None
        for _ in range(3):
            t0 = time.perf_counter(); r = fn(); dt = (time.perf_counter() - t0) * 1e3
            if best is None or dt < best:
                best, ph = dt, [round(x, 3) for x in ll.engine_stats()["last_kernel_ms"][:6]]
        o[name] = {"wall_ms": round(best, 3), "phases_ms": ph}
        if name in res:
            scale = np.max(np.abs(res[name]))
            o[name]["max_abs_diff_over_max"] = float(np.max(np.abs(r - res[name])) / scale)
        else:
            res[name] = r
    out[variant] = o
    for k in kv:
        del os.environ[k]
print(json.dumps(out))
