"""Pure numpy/Python model of the replay plan (tskit_b200/csrc/plan.cuh), used by
the GPU tests to localise a failure to one staging step.  Small inputs only."""
import numpy as np


def build(t, a=None, b=None):
    t.ensure_derived()
    E, N, L = t.num_edges, t.num_nodes, t.sequence_length
    a = 0.0 if a is None else a
    b = L if b is None else b
    el, er, ep, ec = t.edges_left, t.edges_right, t.edges_parent, t.edges_child
    I, O = t.edge_insertion_order, t.edge_removal_order
    tm = t.nodes_time
    # insertion list
    lI = el[I]
    i0 = np.searchsorted(lI, a, side="right")
    i1 = np.searchsorted(lI, b, side="left")
    seeds = I[:i0]
    if a > 0:
        seeds = seeds[er[seeds] > a]
        seeds = seeds[np.argsort(tm[ep[seeds]], kind="stable")]
    ins = np.concatenate([seeds, I[i0:i1]]).astype(np.int64)
    ins_pos = np.maximum(el[ins], a)
    rO = er[O]
    r0 = np.searchsorted(rO, a, side="right")
    r1 = np.searchsorted(rO, b, side="left")
    rem = O[r0:r1].astype(np.int64)
    rem_pos = er[rem]
    n_ins, n_rem = len(ins), len(rem)
    nev = n_ins + n_rem
    idx_ins = np.arange(n_ins) + np.searchsorted(rem_pos, ins_pos, side="right")
    idx_rem = np.arange(n_rem) + np.searchsorted(ins_pos, rem_pos, side="left")
    ev_edge = np.empty(nev, dtype=np.int64)
    ev_pos = np.empty(nev)
    ev_sign = np.empty(nev, dtype=np.int8)
    ev_edge[idx_ins] = ins
    ev_pos[idx_ins] = ins_pos
    ev_sign[idx_ins] = 1
    ev_edge[idx_rem] = rem
    ev_pos[idx_rem] = rem_pos
    ev_sign[idx_rem] = -1
    ev_child = ec[ev_edge]
    ev_parent = ep[ev_edge]
    ev_sbl = ev_sign * (tm[ev_parent] - tm[ev_child])
    # child-major CSR
    order = np.lexsort((el, ec))
    coff = np.searchsorted(ec[order], np.arange(N + 1))
    cl, cr, cp = el[order], er[order], ep[order]

    def span_parent(u, x):
        if not x > a:
            return -1
        lo, hi = coff[u], coff[u + 1]
        k = np.searchsorted(cl[lo:hi], x, side="left")
        if k == 0:
            return -1
        return int(cp[lo + k - 1]) if cr[lo + k - 1] > x else -1

    voff = [0]
    em_node, em_bl, em_ev = [], [], []
    for i in range(nev):
        u = int(ev_parent[i])
        x = ev_pos[i]
        while u != -1:
            v = span_parent(u, x)
            em_node.append(u)
            em_ev.append(i)
            em_bl.append(0.0 if v == -1 else tm[v] - tm[u])
            u = v
        voff.append(len(em_node))
    voff = np.array(voff, dtype=np.uint32)
    em_node = np.array(em_node, dtype=np.int32)
    em_bl = np.array(em_bl, dtype=np.float64)
    em_ev = np.array(em_ev, dtype=np.int64)
    V = len(em_node)
    level = np.zeros(N, dtype=np.uint32)
    o = np.lexsort((ec, ep, tm[ep]))
    for p, c in zip(ep[o].tolist(), ec[o].tolist()):
        if level[c] + 1 > level[p]:
            level[p] = level[c] + 1
    # longest-path levels need children final first: iterate to a fixed point
    changed = True
    while changed:
        nl = level.copy()
        np.maximum.at(nl, ep, level[ec] + 1)
        changed = not np.array_equal(nl, level)
        level = nl
    rank_node = np.lexsort((np.arange(N), level)).astype(np.int32)
    rank = np.empty(N, dtype=np.int64)
    rank[rank_node] = np.arange(N)
    key = rank[em_node] if V else np.zeros(0, dtype=np.int64)
    nm_em = np.argsort(key, kind="stable")
    em_perm = np.empty(V, dtype=np.uint32)
    em_perm[nm_em] = np.arange(V)
    nm_key = key[nm_em].astype(np.uint32)
    nm_ev = em_ev[nm_em]
    noff = np.searchsorted(nm_key, np.arange(N + 1))
    ev_src = np.empty(nev, dtype=np.int32)
    for i in range(nev):
        c = int(ev_child[i])
        lo, hi = noff[rank[c]], noff[rank[c] + 1]
        k = np.searchsorted(nm_ev[lo:hi], i, side="left")
        ev_src[i] = lo + k - 1 if k > 0 else ~c
    nm_src = ev_src[nm_ev] if V else np.zeros(0, dtype=np.int32)
    head = np.ones(V, dtype=bool)
    head[1:] = nm_key[1:] != nm_key[:-1]
    nm_flag = ((ev_sign[nm_ev] < 0).astype(np.uint8) | (head.astype(np.uint8) << 1)) if V \
        else np.zeros(0, dtype=np.uint8)
    nlevels = int(level.max()) + 1 if N else 1
    lvl_sorted = level[rank_node]
    lro = np.searchsorted(lvl_sorted, np.arange(nlevels + 1))
    level_begin = noff[lro].astype(np.uint32)
    return dict(ev_pos=ev_pos, ev_child=ev_child.astype(np.int32), ev_sign=ev_sign,
                ev_sbl=ev_sbl, ev_src=ev_src, voff=voff, em_node=em_node, em_perm=em_perm,
                em_bl=em_bl, nm_src=nm_src.astype(np.int32), nm_flag=nm_flag, nm_key=nm_key,
                level=level, rank_node=rank_node, level_begin=level_begin)


DTYPES = dict(ev_pos=np.float64, ev_child=np.int32, ev_sign=np.int8, ev_sbl=np.float64,
              ev_src=np.int32, voff=np.uint32, em_node=np.int32, em_perm=np.uint32,
              em_bl=np.float64, nm_src=np.int32, nm_flag=np.uint8, nm_key=np.uint32,
              level=np.uint32, rank_node=np.int32, level_begin=np.uint32)
ORDER = ["ev_pos", "ev_child", "ev_sign", "ev_sbl", "voff", "em_node", "em_bl", "level",
         "rank_node", "nm_key", "em_perm", "level_begin", "ev_src", "nm_src", "nm_flag"]


def compare(ll, t, a=None, b=None):
    """Returns the list of plan arrays that differ from the model, in build order."""
    m = build(t, a, b)
    bad = []
    for name in ORDER:
        got = ll.debug_array(name, DTYPES[name])
        if got.shape != m[name].shape or not np.array_equal(got, m[name]):
            bad.append(name)
    return bad
