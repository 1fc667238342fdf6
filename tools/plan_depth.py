"""Analysis tool: dependency depth of the count propagation for a cached workload."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def analyse(t, hist_len=1 << 16):
    so = "/tmp/plan_depth.so"
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", os.path.join(ROOT, "tools", "plan_depth.c"), "-o", so])
    lib = C.CDLL(so)
    t.ensure_derived()
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    out = np.zeros(8, dtype=np.uint64)
    he = np.zeros(hist_len, dtype=np.uint64)
    hn = np.zeros(hist_len, dtype=np.uint64)
    hc = np.zeros(256, dtype=np.uint64)
    lib.plan_depth.argtypes = [C.c_uint64, C.c_uint64, C.c_double] + [C.c_void_p] * 7 + [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64]
    lib.plan_depth(t.num_nodes, t.num_edges, t.sequence_length, p(t.edges_left), p(t.edges_right),
                   p(t.edges_parent), p(t.edges_child), p(t.edge_insertion_order),
                   p(t.edge_removal_order), p(out), p(he), hist_len, p(hn), p(hc), 256)
    return out, he, hn, hc


if __name__ == "__main__":
    import bench
    name = sys.argv[1] if len(sys.argv) > 1 else "small"
    t, W, _ = bench.load_workload(name)
    out, he, hn, hc = analyse(t)
    nev, V, dn, de, pairs, nbp = [int(x) for x in out[:6]]
    print(f'(node,breakpoint) pairs={pairs} breakpoints={nbp} entries per-event layout={V + nev}')
    print(f"nev={nev} V={V} dbar={V / nev:.3f} D_node={dn} D_entry={de}")
    for nm, h, d in (("entry", he, de), ("node", hn, dn)):
        cs = np.cumsum(h[: d + 1])
        print(nm, "levels holding 50/90/99/99.9% of visits:",
              [int(np.searchsorted(cs, q * cs[-1])) for q in (0.5, 0.9, 0.99, 0.999)])
        print(nm, "first 12 level sizes", h[:12].tolist())
        print(nm, "levels with < 1024 visits:", int(np.count_nonzero(h[: d + 1] < 1024)),
              "visits in them:", int(h[: d + 1][h[: d + 1] < 1024].sum()))
    print("chain length hist (0..40):", hc[:41].tolist(), "max", int(np.nonzero(hc)[0].max()))
