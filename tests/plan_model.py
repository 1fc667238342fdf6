"""Pure numpy/Python model of the replay plan (tskit_b200/csrc/plan.cuh), used by
the GPU tests to localise a failure to one staging step.  Small inputs only."""
import os

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
    vis_node = np.array(em_node, dtype=np.int64)
    vis_bl = np.array(em_bl, dtype=np.float64)
    vis_ev = np.array(em_ev, dtype=np.int64)
    V = len(vis_node)
    level = np.zeros(N, dtype=np.uint32)
    changed = True
    while changed:
        nl = level.copy()
        if E:
            np.maximum.at(nl, ep, level[ec] + 1)
        changed = not np.array_equal(nl, level)
        level = nl
    rank_node = np.lexsort((np.arange(N), level)).astype(np.int32)
    rank = np.empty(N, dtype=np.int64)
    rank[rank_node] = np.arange(N)
    # entries, event-major: CHILD entry of event i at voff[i] + i, then its visits
    Ve = V + nev
    em_node = np.zeros(Ve, dtype=np.int64)
    em_ev = np.zeros(Ve, dtype=np.int64)
    em_child = np.zeros(Ve, dtype=bool)
    em_blv = np.zeros(Ve, dtype=np.float64)
    eoff = voff[:-1].astype(np.int64) + np.arange(nev)
    em_node[eoff] = ev_child
    em_ev[eoff] = np.arange(nev)
    em_child[eoff] = True
    em_blv[eoff] = np.where(ev_sign > 0, ev_sbl, 0.0)
    epos = np.arange(V) + vis_ev + 1
    em_node[epos] = vis_node
    em_ev[epos] = vis_ev
    em_blv[epos] = vis_bl
    flag = np.ones(nev, dtype=bool)
    flag[:-1] = ev_pos[:-1] != ev_pos[1:]
    ev_bp = np.concatenate([[0], np.cumsum(flag)[:-1]]).astype(np.int64) if nev else np.zeros(0, dtype=np.int64)
    key = rank[em_node]
    sorted_e = np.argsort(key, kind="stable")
    sorted_key = key[sorted_e]
    noff = np.searchsorted(sorted_key, np.arange(N + 1))
    bp_k = ev_bp[em_ev[sorted_e]]
    end = np.ones(Ve, dtype=bool)
    if Ve > 1:
        end[:-1] = (sorted_key[1:] != sorted_key[:-1]) | (bp_k[1:] != bp_k[:-1])
    endscan = np.concatenate([[0], np.cumsum(end)]).astype(np.int64)
    ends_total = int(endscan[-1])
    inv = np.empty(Ve, dtype=np.int64)
    inv[sorted_e] = np.arange(Ve)
    pidx = endscan[:-1] + sorted_key + 1
    kc = inv[eoff]
    ev_src = pidx[kc] - (ev_sign < 0)
    P = ends_total + N
    i_k = em_ev[sorted_e]
    P_pad = -(-P // 1024) * 1024
    pc_x = np.full(P_pad, -1.0)
    pc_bl = np.zeros(P_pad)
    pc_x[pidx[end]] = ev_pos[i_k[end]]
    pc_bl[pidx[end]] = em_blv[sorted_e[end]]
    piece_rank = np.zeros(P, dtype=np.int64)
    piece_rank[pidx[end]] = sorted_key[end]
    poff = np.concatenate([endscan[noff[:N]] if N else [], [ends_total]]).astype(np.int64) + np.arange(N + 1)
    piece_rank[poff[:N]] = np.arange(N)
    # references of every real piece: own INIT piece if the node is a sample, then the children
    # in the tree right of the piece's breakpoint (parent-major edge list scanned backwards)
    is_sample = (t.nodes_flags & 1).astype(bool)
    pm = np.lexsort((ec, el, ep))  # (parent, left, child)
    pm_off = np.searchsorted(ep[pm], np.arange(N + 1))
    pl, pr, pch = el[pm], er[pm], ec[pm]
    ref_lists = [[] for _ in range(P)]
    for p in range(P):
        x = pc_x[p]
        if x < 0:
            continue
        r = piece_rank[p]
        u = rank_node[r]
        lst = ref_lists[p]
        if is_sample[u]:
            lst.append(int(poff[r]))
        lo, hi = pm_off[u], pm_off[u + 1]
        k = np.searchsorted(pl[lo:hi], x, side="right")
        for j in range(lo + k - 1, lo - 1, -1):
            if pr[j] > x:
                rc = rank[pch[j]]
                q0, q1 = poff[rc], poff[rc + 1]
                lst.append(int(q0 + np.searchsorted(pc_x[q0:q1], x, side="right") - 1))
    height = np.zeros(P, dtype=np.int64)
    changed = True
    while changed:
        changed = False
        for p in range(P):
            h = 0
            for q in ref_lists[p]:
                if pc_x[q] >= 0:
                    h = max(h, height[q] + 1)
            if h != height[p]:
                height[p] = h
                changed = True
    if os.environ.get("TSKB_ORDER", "").startswith("l"):
        height = level[rank_node[piece_rank]].astype(np.int64)
    # pieces that are computed: with a branch above them, or holding the state a mutation reads
    mut_piece = np.zeros(t.num_mutations, dtype=np.int64)
    for m in range(t.num_mutations):
        x = t.sites_position[t.mutations_site[m]]
        r = rank[t.mutations_node[m]]
        lo, hi = poff[r], poff[r + 1]
        mut_piece[m] = lo + np.searchsorted(pc_x[lo:hi], x, side="right") - 1
    needed = (pc_x[:P] >= 0) & (pc_bl[:P] != 0)
    needed[mut_piece[pc_x[mut_piece] >= 0]] = True
    real = np.nonzero(needed)[0]
    if os.environ.get("TSKB_ORDER", "").startswith("x"):
        real = real[np.argsort(pc_x[real], kind="stable")]
    order = real[np.argsort(height[real], kind="stable")]
    nheights = int(height.max()) + 1 if P else 1
    TILE = 1024
    hs = height[order]
    begin = np.searchsorted(hs, np.arange(nheights + 1))
    ntile = -(-(begin[1:] - begin[:-1]) // TILE)
    level_begin = np.concatenate([[0], np.cumsum(ntile) * TILE]).astype(np.uint32)
    npp = int(level_begin[-1])
    # state slots: processing position of every real piece, then one INIT slot per sample, then
    # the zero slot shared by the INIT pieces of nodes that are not samples
    n = int(is_sample.sum())
    sample_index = np.full(N, -1, dtype=np.int64)
    sample_index[np.nonzero(is_sample)[0]] = np.arange(n)
    pos = level_begin[:-1].astype(np.int64)[hs] + (np.arange(len(order)) - begin[hs])
    perm = np.full(P + 1, npp + n, dtype=np.int64)
    perm[order] = pos
    si = sample_index[rank_node]
    perm[poff[:N]] = np.where(si >= 0, npp + si, npp + n)
    # breakpoints: distinct diff positions, then the end of the range
    bp_pos = np.concatenate([np.unique(ev_pos), [b]])
    T = len(bp_pos) - 1
    q_bp0 = np.zeros(npp, dtype=np.uint32)
    q_bp1 = np.full(npp, 0xFFFFFFFF, dtype=np.uint32)
    q_bl = np.zeros(npp)
    cnts = np.zeros(npp + 1, dtype=np.int64)
    q_bp0[pos] = np.searchsorted(bp_pos[:T], pc_x[order])
    nxt = pc_x[np.minimum(order + 1, P_pad - 1)]
    q_bp1[pos] = np.where((order + 1 < P) & (nxt >= 0), np.searchsorted(bp_pos[:T], nxt), T)
    q_bl[pos] = pc_bl[order]
    cnts[pos] = [len(ref_lists[p]) for p in order]
    q_off = np.concatenate([[0], np.cumsum(cnts[:-1])]).astype(np.uint32)
    refs = np.zeros(int(cnts.sum()), dtype=np.uint32)
    for j, p in zip(pos, order):
        refs[q_off[j]:q_off[j] + len(ref_lists[p])] = np.sort(perm[ref_lists[p]])
    # sites
    mut_src = perm[mut_piece].astype(np.int32)
    return dict(ev_pos=ev_pos, ev_child=ev_child.astype(np.int32), ev_sign=ev_sign, voff=voff,
                q_off=q_off, refs=refs, q_bp0=q_bp0, q_bp1=q_bp1, q_bl=q_bl, bp_pos=bp_pos, level=level,
                rank_node=rank_node, level_begin=level_begin, mut_src=mut_src)


DTYPES = dict(ev_pos=np.float64, ev_child=np.int32, ev_sign=np.int8, voff=np.uint32,
              q_off=np.uint32, refs=np.uint32, q_bp0=np.uint32, q_bp1=np.uint32, bp_pos=np.float64,
              q_bl=np.float64, level=np.uint32, rank_node=np.int32, level_begin=np.uint32,
              mut_src=np.int32)
ORDER = ["ev_pos", "ev_child", "ev_sign", "voff", "level", "rank_node", "level_begin", "bp_pos", "q_bp0",
         "q_bp1", "q_bl", "q_off", "refs", "mut_src"]


def compare(ll, t, a=None, b=None):
    """Returns the list of plan arrays that differ from the model, in build order."""
    m = build(t, a, b)
    bad = []
    for name in ORDER:
        got = ll.debug_array(name, DTYPES[name])
        if name == "refs":
            got = got[:len(m[name])]  # the device array is over-allocated
        if got.shape != m[name].shape or not np.array_equal(got, m[name]):
            bad.append(name)
    return bad
