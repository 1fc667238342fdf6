"""Analysis tool: structural no-op pieces of a cached workload (tools/noop_stats.cpp)."""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
so = "/tmp/noop_stats.so"
subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", os.path.join(ROOT, "tools", "noop_stats.cpp"), "-o", so])
lib = C.CDLL(so)
name = sys.argv[1] if len(sys.argv) > 1 else "small"
t, W, _ = bench.load_workload(name)
t.ensure_derived()
p = lambda a: a.ctypes.data_as(C.c_void_p)
out = np.zeros(8, dtype=np.uint64)
lib.noop_stats.argtypes = [C.c_uint64, C.c_uint64, C.c_double] + [C.c_void_p] * 9
lib.noop_stats(t.num_nodes, t.num_edges, t.sequence_length, p(t.edges_left), p(t.edges_right), p(t.edges_parent),
               p(t.edges_child), p(t.edge_insertion_order), p(t.edge_removal_order), p(t.nodes_time),
               p(t.nodes_flags), p(out))
pieces, sn, fn, nbp, roots, first = [int(x) for x in out[:6]]
print(f"pieces {pieces} breakpoints {nbp} first-pieces {first} root pieces {roots / pieces:.3f}")
print(f"state no-ops {sn / pieces:.3f} (of which parent unchanged too: {fn / pieces:.3f})")
