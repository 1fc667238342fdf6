"""ctypes binding of oracle/stats_oracle.c (the plain-C restatement).  TEST
INFRASTRUCTURE ONLY -- see the header of stats_oracle.c."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "stats_oracle.c")
_OUT = os.path.join(_HERE, "_build", "liboracle.so")

STAT_IDS = {"diversity": 0, "segregating_sites": 1, "Y1": 2, "divergence": 3, "Y2": 4,
            "f2": 5, "genetic_relatedness": 6, "Y3": 7, "f3": 8, "f4": 9}
TUPLE = {"divergence": 2, "Y2": 2, "f2": 2, "genetic_relatedness": 2, "Y3": 3, "f3": 3,
         "f4": 4}
SUMMARY_FUNC = C.CFUNCTYPE(C.c_int, C.c_uint64, C.POINTER(C.c_double), C.c_uint64,
                           C.POINTER(C.c_double), C.c_void_p)


def build(force=False):
    if force or not os.path.exists(_OUT) or os.path.getmtime(_OUT) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_OUT), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-Wall", "-shared", "-fPIC", _SRC,
                               "-lm", "-o", _OUT])
    return _OUT


class OrcTables(C.Structure):
    _fields_ = [
        ("sequence_length", C.c_double), ("time_uncalibrated", C.c_int32),
        ("num_nodes", C.c_uint64), ("node_flags", C.c_void_p), ("node_time", C.c_void_p),
        ("num_edges", C.c_uint64), ("edge_left", C.c_void_p), ("edge_right", C.c_void_p),
        ("edge_parent", C.c_void_p), ("edge_child", C.c_void_p),
        ("edge_insertion_order", C.c_void_p), ("edge_removal_order", C.c_void_p),
        ("num_sites", C.c_uint64), ("site_position", C.c_void_p),
        ("site_ancestral_state", C.c_void_p), ("site_ancestral_state_offset", C.c_void_p),
        ("num_mutations", C.c_uint64), ("mutation_site", C.c_void_p),
        ("mutation_node", C.c_void_p), ("mutation_parent", C.c_void_p),
        ("mutation_derived_state", C.c_void_p), ("mutation_derived_state_offset", C.c_void_p)]


class OracleError(Exception):
    def __init__(self, code):
        self.code = code
        super().__init__(f"oracle error {code}")


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _flags(mode, span_normalise=True, polarised=False, centre=True):
    f = {"site": 1, "branch": 2, "node": 4, None: 1}[mode]
    if span_normalise:
        f |= 1 << 11
    if polarised:
        f |= 1 << 10
    if not centre:
        f |= 1 << 14
    return f


class Oracle:
    def __init__(self, tables):
        self.lib = C.CDLL(build())
        tables.ensure_derived()
        self.t = tables
        o = OrcTables()
        o.sequence_length = tables.sequence_length
        o.time_uncalibrated = int(tables.time_uncalibrated)
        o.num_nodes, o.node_flags, o.node_time = tables.num_nodes, _p(tables.nodes_flags), _p(tables.nodes_time)
        o.num_edges = tables.num_edges
        o.edge_left, o.edge_right = _p(tables.edges_left), _p(tables.edges_right)
        o.edge_parent, o.edge_child = _p(tables.edges_parent), _p(tables.edges_child)
        o.edge_insertion_order = _p(tables.edge_insertion_order)
        o.edge_removal_order = _p(tables.edge_removal_order)
        o.num_sites, o.site_position = tables.num_sites, _p(tables.sites_position)
        o.site_ancestral_state = _p(tables.sites_ancestral_state)
        o.site_ancestral_state_offset = _p(tables.sites_ancestral_state_offset)
        o.num_mutations = tables.num_mutations
        o.mutation_site, o.mutation_node = _p(tables.mutations_site), _p(tables.mutations_node)
        o.mutation_parent = _p(tables.mutations_parent)
        o.mutation_derived_state = _p(tables.mutations_derived_state)
        o.mutation_derived_state_offset = _p(tables.mutations_derived_state_offset)
        self.o = o

    def _windows(self, windows):
        if windows is None:
            windows = [0.0, self.t.sequence_length]
        return np.ascontiguousarray(windows, dtype=np.float64)

    @staticmethod
    def _sets(sample_sets):
        sizes = np.array([len(s) for s in sample_sets], dtype=np.uint64)
        flat = (np.concatenate([np.asarray(s, dtype=np.int32) for s in sample_sets])
                if len(sample_sets) else np.zeros(0, dtype=np.int32))
        return sizes, np.ascontiguousarray(flat, dtype=np.int32)

    def stat(self, name, sample_sets, indexes=None, windows=None, mode="site",
             span_normalise=True, polarised=False, centre=True):
        sizes, flat = self._sets(sample_sets)
        w = self._windows(windows)
        if name in TUPLE:
            idx = np.ascontiguousarray(indexes, dtype=np.int32).reshape(-1, TUPLE[name])
            M = len(idx)
        else:
            idx, M = None, len(sizes)
        res = np.empty((len(w) - 1, max(M, 1)), dtype=np.float64)
        ret = self.lib.orc_sample_count_stat(
            C.byref(self.o), C.c_int(STAT_IDS[name]), C.c_uint64(len(sizes)), _p(sizes),
            _p(flat), C.c_uint64(0 if idx is None else len(idx)), _p(idx),
            C.c_uint64(len(w) - 1), _p(w),
            C.c_uint32(_flags(mode, span_normalise, polarised, centre)), _p(res))
        if ret != 0:
            raise OracleError(ret)
        return res[:, :M]

    def general_stat(self, weights, f, output_dim, windows=None, mode="site",
                     span_normalise=True, polarised=False):
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        w = self._windows(windows)
        res = np.empty((len(w) - 1, output_dim), dtype=np.float64)

        def tramp(k, state, m, result, params):
            y = np.asarray(f(np.ctypeslib.as_array(state, shape=(k,)).copy()), dtype=np.float64)
            for j in range(m):
                result[j] = y[j]
            return 0

        cb = SUMMARY_FUNC(tramp)
        ret = self.lib.orc_general_stat(
            C.byref(self.o), C.c_uint64(weights.shape[1]), _p(weights), C.c_uint64(output_dim),
            cb, None, C.c_uint64(len(w) - 1), _p(w),
            C.c_uint32(_flags(mode, span_normalise, polarised)), _p(res))
        if ret != 0:
            raise OracleError(ret)
        return res

    # ---- weighted statistics: the reference's weight pre-processing and summary functions
    # restated in numpy on top of general_stat (pinned against the reference package in
    # tests/test_oracle.py)
    def trait_covariance(self, W, windows=None, mode="site", span_normalise=True):
        """tsk_treeseq_trait_covariance, c/tskit/trees.c:3960-4022."""
        W = np.asarray(W, dtype=np.float64)
        n = self.t.num_samples
        Wc = W - W.sum(axis=0) / n
        return self.general_stat(Wc, lambda x: (x * x) / (2 * (n - 1) * (n - 1)), W.shape[1],
                                 windows=windows, mode=mode, span_normalise=span_normalise)

    def trait_correlation(self, W, windows=None, mode="site", span_normalise=True):
        """tsk_treeseq_trait_correlation, c/tskit/trees.c:4024-4110."""
        W = np.asarray(W, dtype=np.float64)
        n, K = self.t.num_samples, W.shape[1]
        means = W.sum(axis=0) / n
        meansqs = ((W * W).sum(axis=0) - means * means * n) / (n - 1)
        Ws = np.column_stack([(W - means) / np.sqrt(meansqs), np.full(n, 1.0 / n)])

        def f(x):
            p = x[K]
            if 0.0 < p < 1.0:
                return (x[:K] * x[:K]) / (2 * (p * (1 - p)) * n * (n - 1))
            return np.zeros(K)
        return self.general_stat(Ws, f, K, windows=windows, mode=mode, span_normalise=span_normalise)

    def trait_linear_model(self, W, Z, windows=None, mode="site", span_normalise=True):
        """tsk_treeseq_trait_linear_model, c/tskit/trees.c:4106-4219 (Z already orthonormalised, as
        the low-level reference function assumes)."""
        W = np.asarray(W, dtype=np.float64)
        Z = np.asarray(Z, dtype=np.float64)
        n, K, Cn = self.t.num_samples, W.shape[1], Z.shape[1]
        V = W.T @ Z
        Wn = np.column_stack([W, Z, np.ones(n)])

        def f(x):
            m = x[K + Cn]
            out = np.zeros(K)
            if 0.0 < m < n:
                z = x[K:K + Cn]
                for i in range(K):
                    a = x[i] - (z * V[i]).sum()
                    denom = m - (z * z).sum()
                    out[i] = 0.0 if denom < 1e-8 else (a * a) / (2 * denom * denom)
            return out
        return self.general_stat(Wn, f, K, windows=windows, mode=mode, span_normalise=span_normalise)

    def genetic_relatedness_weighted(self, W, indexes, windows=None, mode="site", span_normalise=True,
                                     polarised=False, centre=True):
        """tsk_treeseq_genetic_relatedness_weighted, c/tskit/trees.c:4800-4897."""
        W = np.asarray(W, dtype=np.float64)
        n, K = self.t.num_samples, W.shape[1]
        idx = np.asarray(indexes, dtype=np.int64).reshape(-1, 2)
        Wn = np.column_stack([W, np.full(n, 1.0 / n)])
        tot = np.concatenate([W.sum(axis=0), [1.0]])

        def f(x):
            if centre:
                pn = x[K]
                return (x[idx[:, 0]] - tot[idx[:, 0]] * pn) * (x[idx[:, 1]] - tot[idx[:, 1]] * pn)
            return x[idx[:, 0]] * x[idx[:, 1]]
        return self.general_stat(Wn, f, len(idx), windows=windows, mode=mode,
                                 span_normalise=span_normalise, polarised=polarised)

    def site_allele_frequency_spectrum(self, sample_sets, windows=None, span_normalise=True,
                                       polarised=False):
        """tsk_treeseq_allele_frequency_spectrum in site mode (c/tskit/trees.c:3469-3648, 3814-3928),
        by definition: per site and allele carried by some but not all samples, +1 (polarised,
        ancestral allele skipped) or +1/2 (folded, :3469-3495) at the vector of per-set allele
        counts.  Pinned against the reference package in tests/test_dropin.py."""
        t = self.t
        w = self._windows(windows)
        W = len(w) - 1
        sets = [np.asarray(x, dtype=np.int64) for x in sample_sets]
        dims = [len(x) + 1 for x in sets]
        out = np.zeros([W] + dims)
        n_all = t.num_samples

        def fold(coord):
            n = sum(d - 1 for d in dims) / 2
            s = int(sum(coord))
            k = len(dims)
            while s == n and k > 0:
                k -= 1
                n -= (dims[k] - 1) / 2
                s -= int(coord[k])
            if s > n:
                return tuple(dims[k] - 1 - int(coord[k]) for k in range(len(dims)))
            return tuple(int(c) for c in coord)

        G = self.genotype_matrix(isolated_as_missing=False)
        col = {int(u): i for i, u in enumerate(t.samples)}
        idx = [np.array([col[int(u)] for u in x], dtype=np.int64) for x in sets]
        inc = 1.0 if polarised else 0.5
        for j in range(t.num_sites):
            win = int(np.searchsorted(w, t.sites_position[j], side="right")) - 1
            g = G[j]
            for a in range(1 if polarised else 0, int(g.max()) + 1):
                tot = int((g == a).sum())
                if 0 < tot < n_all:
                    coord = [int((g[i] == a).sum()) for i in idx]
                    out[(win,) + (tuple(coord) if polarised else fold(coord))] += inc
        if span_normalise:
            out /= (w[1:] - w[:-1]).reshape([W] + [1] * len(dims))
        return out

    def genetic_relatedness_vector(self, W, windows=None, nodes=None, centre=True, span_normalise=True):
        """tsk_treeseq_genetic_relatedness_vector (c/tskit/trees.c:10445-10816) restated step by step:
        the matvec calculator's per-node vectors w (summed weights below), v (accumulated
        branch area x w, handed down on edge removal) and x (position of the last update), the edge
        diffs between windows[0] and windows[-1], and the output pass over the focal nodes' ancestors.
        Small inputs only.  Pinned against the reference package in tests/test_dropin.py."""
        t = self.t
        win = self._windows(windows)
        nw = len(win) - 1
        W = np.asarray(W, dtype=np.float64)
        if W.ndim == 1:
            W = W.reshape(-1, 1)
        samples = t.samples
        n, K = len(samples), W.shape[1]
        focal = samples if nodes is None else np.asarray(nodes, dtype=np.int64)
        N, E = t.num_nodes, t.num_edges
        time = t.nodes_time
        el, er, ep, ec = t.edges_left, t.edges_right, t.edges_parent, t.edges_child
        I, O = t.edge_insertion_order, t.edge_removal_order
        parent = np.full(N, -1, dtype=np.int64)
        x = np.zeros(N)
        v = np.zeros((N, K))
        w = np.zeros((N, K))
        means = np.zeros(K)
        if centre:  # trees.c:10512-10523
            for j in range(n):
                means += W[j]
            means /= n
        for j in range(n):
            w[samples[j]] = W[j] - means
        result = np.zeros((nw, len(focal), K))
        state = {"pos": win[0]}

        def add_z(u):  # trees.c:10552-10571
            p = parent[u]
            if p != -1:
                v[u] += (time[p] - time[u]) * (state["pos"] - x[u]) * w[u]
            x[u] = state["pos"]

        def adjust_path_up(p, c, sign):  # trees.c:10573-10608
            while p != -1:
                add_z(p)
                v[c] -= sign * v[p]
                w[p] += sign * w[c]
                p = parent[p]

        def remove_edge(p, c):  # trees.c:10610-10625
            add_z(c)
            parent[c] = -1
            adjust_path_up(p, c, -1)

        def insert_edge(p, c):  # trees.c:10627-10635
            adjust_path_up(p, c, +1)
            x[c] = state["pos"]
            parent[c] = p

        def write_output(m):  # trees.c:10637-10699
            for j, u in enumerate(focal):
                u = int(u)
                while u != -1:
                    if x[u] != state["pos"]:
                        add_z(u)
                    result[m, j] += v[u]
                    u = parent[u]
            if centre:
                om = np.zeros(K)
                for j in range(len(focal)):
                    om += result[m, j]
                om /= len(focal)
                result[m] -= om
            v[:] = 0

        # the tree holding windows[0]: every edge with left <= windows[0] < right (trees.c:10733-10741)
        j = 0
        while j < E and el[I[j]] <= win[0]:
            e = I[j]
            if win[0] < er[e]:
                insert_edge(int(ep[e]), int(ec[e]))
            j += 1
        k = 0
        while k < E and er[O[k]] <= win[0]:
            k += 1
        m = 0
        while m < nw:  # trees.c:10746-10782
            pos = state["pos"]
            while k < E and er[O[k]] == pos:
                e = O[k]
                remove_edge(int(ep[e]), int(ec[e]))
                k += 1
            while j < E and el[I[j]] == pos:
                e = I[j]
                insert_edge(int(ep[e]), int(ec[e]))
                j += 1
            nxt = win[m + 1]
            if j < E:
                nxt = min(nxt, el[I[j]])
            if k < E:
                nxt = min(nxt, er[O[k]])
            state["pos"] = nxt
            if nxt == win[m + 1]:
                write_output(m)
                m += 1
        if span_normalise:
            result /= (win[1:] - win[:-1]).reshape(nw, 1, 1)
        return result

    def branch_allele_frequency_spectrum(self, sample_sets, windows=None, span_normalise=True,
                                         polarised=False):
        """tsk_treeseq_branch_allele_frequency_spectrum + tsk_treeseq_update_branch_afs
        (c/tskit/trees.c:3650-3812) restated step by step for the default time window [0, inf),
        including the reference's bookkeeping of `last_update` (a node is credited from the last time
        it was visited or flushed, which is earlier than the insertion of its edge when it was
        parentless before).  Small inputs only.  Pinned against the reference package in
        tests/test_dropin.py."""
        t = self.t
        w = self._windows(windows)
        W = len(w) - 1
        sets = [np.asarray(x, dtype=np.int64) for x in sample_sets]
        K = len(sets)
        dims = [len(x) + 1 for x in sets]
        out = np.zeros([W] + dims)
        N, n_all, L = t.num_nodes, t.num_samples, t.sequence_length
        time = t.nodes_time
        count = np.zeros((N, K + 1), dtype=np.int64)
        for k, x in enumerate(sets):
            count[x, k] = 1
        count[t.samples, K] = 1
        parent = np.full(N, -1, dtype=np.int64)
        last_update = np.zeros(N)
        el, er, ep, ec = t.edges_left, t.edges_right, t.edges_parent, t.edges_child
        I, O = t.edge_insertion_order, t.edge_removal_order
        E = t.num_edges

        def fold(coord):
            n = sum(d - 1 for d in dims) / 2
            s = int(sum(coord))
            k = len(dims)
            while s == n and k > 0:
                k -= 1
                n -= (dims[k] - 1) / 2
                s -= int(coord[k])
            if s > n:
                return tuple(dims[k] - 1 - int(coord[k]) for k in range(len(dims)))
            return tuple(int(c) for c in coord)

        def update(u, right, wi):
            if parent[u] != -1:
                t_u, t_v = time[u], time[parent[u]]
                if 0 < count[u, K] < n_all and 0.0 < t_v:
                    c = count[u, :K]
                    c = tuple(int(v) for v in c) if polarised else fold(c)
                    blen = max(0.0, min(np.inf, t_v) - max(0.0, t_u))
                    out[(wi,) + c] += (right - last_update[u]) * blen
            last_update[u] = right

        tj = tk = 0
        t_left = 0.0
        wi = 0
        while tj < E or t_left < L:
            while tk < E and er[O[tk]] == t_left:
                h = O[tk]
                tk += 1
                u, v = int(ec[h]), int(ep[h])
                update(u, t_left, wi)
                while v != -1:
                    update(v, t_left, wi)
                    count[v] -= count[u]
                    v = int(parent[v])
                parent[u] = -1
            while tj < E and el[I[tj]] == t_left:
                h = I[tj]
                tj += 1
                u, v = int(ec[h]), int(ep[h])
                parent[u] = v
                while v != -1:
                    update(v, t_left, wi)
                    count[v] += count[u]
                    v = int(parent[v])
            t_right = L
            if tj < E:
                t_right = min(t_right, el[I[tj]])
            if tk < E:
                t_right = min(t_right, er[O[tk]])
            while wi < W and w[wi + 1] <= t_right:
                for u in range(N):
                    update(u, w[wi + 1], wi)
                wi += 1
            t_left = t_right
        if span_normalise:
            out /= (w[1:] - w[:-1]).reshape([W] + [1] * len(dims))
        return out

    def trees_at(self, positions, tracked=None):
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        order = np.argsort(pos, kind="stable")
        N = self.t.num_nodes
        par = np.empty((len(pos), N), dtype=np.int32)
        cnt = np.empty((len(pos), N), dtype=np.int32)
        tr = None if tracked is None else np.ascontiguousarray(tracked, dtype=np.int32)
        sp = np.ascontiguousarray(pos[order])
        ret = self.lib.orc_trees_at(C.byref(self.o), C.c_uint64(len(pos)), _p(sp), _p(tr),
                                    C.c_uint64(0 if tr is None else len(tr)), _p(par), _p(cnt))
        if ret != 0:
            raise OracleError(ret)
        inv = np.empty_like(order)
        inv[order] = np.arange(len(order))
        return par[inv], cnt[inv]

    def genotype_matrix(self, samples=None, isolated_as_missing=True):
        s = self.t.samples if samples is None else np.ascontiguousarray(samples, dtype=np.int32)
        out = np.empty((self.t.num_sites, len(s)), dtype=np.int32)
        ret = self.lib.orc_genotype_matrix(C.byref(self.o), _p(s), C.c_uint64(len(s)),
                                           C.c_uint32(0 if isolated_as_missing else 2), _p(out))
        if ret != 0:
            raise OracleError(ret)
        return out

    def divergence_matrix(self, sample_sets=None, windows=None, mode="site",
                          span_normalise=True):
        w = self._windows(windows)
        if sample_sets is None:
            sizes = flat = None
            n = self.t.num_samples
        else:
            sizes, flat = self._sets(sample_sets)
            n = len(sizes)
        res = np.empty((len(w) - 1, n, n), dtype=np.float64)
        ret = self.lib.orc_divergence_matrix(
            C.byref(self.o), C.c_uint64(n), _p(sizes), _p(flat), C.c_uint64(len(w) - 1), _p(w),
            C.c_uint32(_flags(mode, span_normalise)), _p(res))
        if ret != 0:
            raise OracleError(ret)
        return res
