"""Seeded synthetic inputs: infinite-sites mutation placement here;
the Wright-Fisher ancestry generator lives in tskit_b200/csrc/wfsim.cpp."""
import numpy as np

from .tables import Tables


def add_mutations(t: Tables, target, seed=1):
    """Infinite-sites mutations: `target` draws of an edge with probability
    proportional to span x branch length, position = floor(uniform within the
    edge's span), duplicate positions dropped (one mutation per site, ancestral
    "0", derived "1")."""
    rng = np.random.default_rng(seed)
    span = t.edges_right - t.edges_left
    bl = t.nodes_time[t.edges_parent] - t.nodes_time[t.edges_child]
    wgt = span * bl
    cdf = np.cumsum(wgt)
    e = np.searchsorted(cdf, rng.random(target) * cdf[-1], side="right")
    e = np.minimum(e, t.num_edges - 1)
    pos = np.floor(t.edges_left[e] + rng.random(target) * span[e])
    pos = np.minimum(pos, np.nextafter(t.edges_right[e], -np.inf))
    pos = np.maximum(pos, t.edges_left[e])
    pos, first = np.unique(pos, return_index=True)
    e = e[first]
    S = len(pos)
    t.sites_position = pos.astype(np.float64)
    t.sites_ancestral_state = np.full(S, ord("0"), dtype=np.int8)
    t.sites_ancestral_state_offset = np.arange(S + 1, dtype=np.uint64)
    t.mutations_site = np.arange(S, dtype=np.int32)
    t.mutations_node = t.edges_child[e].astype(np.int32)
    t.mutations_parent = np.full(S, -1, dtype=np.int32)
    t.mutations_derived_state = np.full(S, ord("1"), dtype=np.int8)
    t.mutations_derived_state_offset = np.arange(S + 1, dtype=np.uint64)
    return t


def wright_fisher(n, generations, L, ncross=1, seed=42, num_threads=0):
    """Seeded haploid Wright-Fisher ARG of `n` samples (population size n,
    `generations` generations, `ncross` crossovers per meiosis at integer
    positions in [1, L-1]), already simplified; see csrc/wfsim.cpp."""
    import ctypes as C
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtskb_sim.so")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: run `make -C tskit_b200/csrc`")
    lib = C.CDLL(path)
    lib.tskb_wfsim_run.restype = C.c_void_p
    lib.tskb_wfsim_run.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_uint32, C.c_uint64,
                                   C.c_uint32]
    lib.tskb_wfsim_sizes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.tskb_wfsim_copy.argtypes = [C.c_void_p] * 9
    lib.tskb_wfsim_free.argtypes = [C.c_void_p]
    h = lib.tskb_wfsim_run(n, generations, float(L), ncross, seed, num_threads)
    try:
        N, E = C.c_uint64(), C.c_uint64()
        lib.tskb_wfsim_sizes(h, C.byref(N), C.byref(E))
        N, E = N.value, E.value
        flags = np.empty(N, dtype=np.uint32)
        time = np.empty(N, dtype=np.float64)
        left = np.empty(E, dtype=np.float64)
        right = np.empty(E, dtype=np.float64)
        parent = np.empty(E, dtype=np.int32)
        child = np.empty(E, dtype=np.int32)
        ins = np.empty(E, dtype=np.int32)
        rem = np.empty(E, dtype=np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        lib.tskb_wfsim_copy(h, p(flags), p(time), p(left), p(right), p(parent), p(child),
                            p(ins), p(rem))
    finally:
        lib.tskb_wfsim_free(h)
    return Tables(float(L), flags, time, left, right, parent, child,
                  edge_insertion_order=ins, edge_removal_order=rem)
