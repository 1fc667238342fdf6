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
    windows."""
    t.ensure_derived()
    if copies <= 1:
        return t
    return concat_genomes([t] * copies)


def concat_genomes(parts) -> Tables:
    """One tree sequence whose genome is the ``parts`` laid end to end (part r over
    ``[sum of the lengths before it, ...)``): unlinked chromosomes of the same samples.  Every part
    must have the same sample nodes (same flags and times, in the same order among the samples) and
    no sample with children; ancestors are private to their part.  The result is in canonical order
    with its edge indexes built from the parts' indexes (no global sort of the edges)."""
    parts = [p.ensure_derived() for p in parts]
    first = parts[0]
    is_s0 = (first.nodes_flags & 1).astype(bool)
    ns = int(is_s0.sum())
    flags = [first.nodes_flags[is_s0]]
    time = [first.nodes_time[is_s0]]
    left, right, parent, child, ins, rem = [], [], [], [], [], []
    sites = {k: [] for k in ("pos", "anc", "anc_off", "msite", "mnode", "mpar", "der", "der_off")}
    x0, n_inner, e0, s0, m0, a0, d0 = 0.0, 0, 0, 0, 0, 0, 0
    for t in parts:
        is_s = (t.nodes_flags & 1).astype(bool)
        if int(is_s.sum()) != ns or not np.array_equal(t.nodes_time[is_s], time[0]):
            raise ValueError("concat_genomes needs the same samples in every part")
        if t.num_edges and is_s[t.edges_parent].any():
            raise ValueError("concat_genomes needs sample nodes without children")
        shared = np.cumsum(is_s) - 1
        inner = ns + n_inner + np.cumsum(~is_s) - 1
        node_map = np.where(is_s, shared, inner).astype(np.int32)
        flags.append(t.nodes_flags[~is_s])
        time.append(t.nodes_time[~is_s])
        left.append(t.edges_left + x0)
        right.append(t.edges_right + x0)
        parent.append(node_map[t.edges_parent])
        child.append(node_map[t.edges_child])
        ins.append(t.edge_insertion_order.astype(np.int64) + e0)
        rem.append(t.edge_removal_order.astype(np.int64) + e0)
        if t.num_sites:
            sites["pos"].append(t.sites_position + x0)
            sites["anc"].append(t.sites_ancestral_state)
            sites["anc_off"].append(t.sites_ancestral_state_offset[:-1] + np.uint64(a0))
            sites["msite"].append(t.mutations_site + s0)
            sites["mnode"].append(node_map[t.mutations_node])
            sites["mpar"].append(np.where(t.mutations_parent >= 0, t.mutations_parent + m0, -1))
            sites["der"].append(t.mutations_derived_state)
            sites["der_off"].append(t.mutations_derived_state_offset[:-1] + np.uint64(d0))
            a0 += int(t.sites_ancestral_state_offset[-1])
            d0 += int(t.mutations_derived_state_offset[-1])
        x0 += float(t.sequence_length)
        n_inner += int((~is_s).sum())
        e0 += t.num_edges
        s0 += t.num_sites
        m0 += t.num_mutations
    flags, time = np.concatenate(flags), np.concatenate(time)
    parent = np.concatenate(parent)
    # canonical order (time[parent], parent, child, left): every part is already sorted and parent
    # ids grow with the part, so a stable sort on the parent's time is enough
    perm = np.argsort(time[parent], kind="stable")
    inv = np.empty(len(perm), dtype=np.int64)
    inv[perm] = np.arange(len(perm))
    kw = {}
    if s0:
        cat = np.concatenate
        kw = dict(sites_position=cat(sites["pos"]), sites_ancestral_state=cat(sites["anc"]),
                  sites_ancestral_state_offset=cat(sites["anc_off"] + [np.array([a0], dtype=np.uint64)]),
                  mutations_site=cat(sites["msite"]), mutations_node=cat(sites["mnode"]),
                  mutations_parent=cat(sites["mpar"]), mutations_derived_state=cat(sites["der"]),
                  mutations_derived_state_offset=cat(sites["der_off"] + [np.array([d0], dtype=np.uint64)]))
    return Tables(x0, flags, time, np.concatenate(left)[perm], np.concatenate(right)[perm], parent[perm],
                  np.concatenate(child)[perm], time_uncalibrated=first.time_uncalibrated,
                  edge_insertion_order=inv[np.concatenate(ins)].astype(np.int32),
                  edge_removal_order=inv[np.concatenate(rem)].astype(np.int32), **kw)
