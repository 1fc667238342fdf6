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


def repeat_genome(t: Tables, copies: int) -> Tables:
    """One tree sequence of length ``copies * L``: the ARG of ``t`` laid end to end ``copies`` times
    (copy r over ``[r L, (r + 1) L)``), all copies over the same sample nodes, every copy with its
    own ancestors -- a genome of ``copies`` unlinked, identically distributed chromosomes.  Used to
    grow a workload along the genome (weak scaling over genome shards) at the cost of one
    simulation; any window statistic over copy r equals the statistic of ``t`` over the same
    windows.  The result is in canonical order with its edge indexes built."""
    t.ensure_derived()
    if copies <= 1:
        return t
    L = float(t.sequence_length)
    N0, E0, S0, M0 = t.num_nodes, t.num_edges, t.num_sites, t.num_mutations
    is_s = (t.nodes_flags & 1).astype(bool)
    ns = int(is_s.sum())
    ni = N0 - ns
    if is_s[t.edges_parent].any():
        raise ValueError("repeat_genome needs sample nodes without children")
    # node ids: samples first (shared by all copies), then the ancestors of copy 0, copy 1, ...
    shared = np.cumsum(is_s) - 1
    inner = ns + np.cumsum(~is_s) - 1
    flags = np.concatenate([t.nodes_flags[is_s]] + [t.nodes_flags[~is_s]] * copies)
    time = np.concatenate([t.nodes_time[is_s]] + [t.nodes_time[~is_s]] * copies)

    def node_map(u, r):
        return np.where(is_s[u], shared[u], inner[u] + r * ni).astype(np.int32)

    left = np.concatenate([t.edges_left + r * L for r in range(copies)])
    right = np.concatenate([t.edges_right + r * L for r in range(copies)])
    parent = np.concatenate([node_map(t.edges_parent, r) for r in range(copies)])
    child = np.concatenate([node_map(t.edges_child, r) for r in range(copies)])
    # canonical order (time[parent], parent, child, left): every copy is already sorted and parent
    # ids grow with the copy, so a stable sort on the parent's time is enough
    perm = np.argsort(time[parent], kind="stable")
    inv = np.empty(len(perm), dtype=np.int64)
    inv[perm] = np.arange(len(perm))
    ins = inv[np.concatenate([t.edge_insertion_order.astype(np.int64) + r * E0 for r in range(copies)])]
    rem = inv[np.concatenate([t.edge_removal_order.astype(np.int64) + r * E0 for r in range(copies)])]
    kw = {}
    if S0:
        kw["sites_position"] = np.concatenate([t.sites_position + r * L for r in range(copies)])
        kw["sites_ancestral_state"] = np.tile(t.sites_ancestral_state, copies)
        a_len = int(t.sites_ancestral_state_offset[-1])
        kw["sites_ancestral_state_offset"] = np.concatenate(
            [t.sites_ancestral_state_offset[:-1] + np.uint64(r * a_len) for r in range(copies)]
            + [np.array([copies * a_len], dtype=np.uint64)])
        kw["mutations_site"] = np.concatenate([t.mutations_site + r * S0 for r in range(copies)])
        kw["mutations_node"] = np.concatenate([node_map(t.mutations_node, r) for r in range(copies)])
        kw["mutations_parent"] = np.concatenate(
            [np.where(t.mutations_parent >= 0, t.mutations_parent + r * M0, -1) for r in range(copies)])
        kw["mutations_derived_state"] = np.tile(t.mutations_derived_state, copies)
        d_len = int(t.mutations_derived_state_offset[-1])
        kw["mutations_derived_state_offset"] = np.concatenate(
            [t.mutations_derived_state_offset[:-1] + np.uint64(r * d_len) for r in range(copies)]
            + [np.array([copies * d_len], dtype=np.uint64)])
    return Tables(copies * L, flags, time, left[perm], right[perm], parent[perm], child[perm],
                  time_uncalibrated=t.time_uncalibrated,
                  edge_insertion_order=ins.astype(np.int32), edge_removal_order=rem.astype(np.int32),
                  **kw)
