"""ctypes binding of oracle/_ref/libtskit_ref.so (the UNMODIFIED reference C
library behind oracle/ref_shim.c).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never by anything under tskit_b200/.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libtskit_ref.so")

STAT_SITE, STAT_BRANCH, STAT_NODE = 1, 2, 4
STAT_POLARISED, STAT_SPAN_NORMALISE = 1 << 10, 1 << 11
STAT_ALLOW_TIME_UNCALIBRATED, STAT_PAIR_NORMALISE, STAT_NONCENTRED = 1 << 12, 1 << 13, 1 << 14
ISOLATED_NOT_MISSING = 1 << 1

ONE_WAY = {"diversity": 0, "segregating_sites": 1, "Y1": 2}
K_WAY = {"divergence": 0, "Y2": 1, "f2": 2, "genetic_relatedness": 3, "Y3": 4,
         "f3": 5, "f4": 6}
TUPLE = {"divergence": 2, "Y2": 2, "f2": 2, "genetic_relatedness": 2, "Y3": 3,
         "f3": 3, "f4": 4}

GENERAL_STAT_FUNC = C.CFUNCTYPE(C.c_int, C.c_uint64, C.POINTER(C.c_double), C.c_uint64,
                                C.POINTER(C.c_double), C.c_void_p)


def available():
    return os.path.exists(_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(
                f"{_PATH} missing: run `sh oracle/build_ref.sh` where /root/reference exists")
        _lib = C.CDLL(_PATH)
        _lib.ref_strerror.restype = C.c_char_p
        _lib.ref_num_trees.restype = C.c_uint64
        _lib.ref_num_samples.restype = C.c_uint64
    return _lib


class RefError(Exception):
    def __init__(self, code):
        self.code = code
        super().__init__(f"{lib().ref_strerror(C.c_int(code)).decode()} ({code})")


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _check(ret):
    if ret != 0:
        raise RefError(ret)


def mode_flags(mode="site", span_normalise=True, polarised=False, centre=True):
    f = {"site": STAT_SITE, "branch": STAT_BRANCH, "node": STAT_NODE, None: STAT_SITE}[mode]
    if span_normalise:
        f |= STAT_SPAN_NORMALISE
    if polarised:
        f |= STAT_POLARISED
    if not centre:
        f |= STAT_NONCENTRED
    return f


class RefTreeSequence:
    """The reference `tsk_treeseq_t` built from a tskit_b200.tables.Tables."""

    def __init__(self, tables):
        L = lib()
        self.t = tables
        h = C.c_void_p()
        mp = tables.mutations_parent
        _check(L.ref_treeseq_new(
            C.byref(h), C.c_double(tables.sequence_length), C.c_int(tables.time_uncalibrated),
            C.c_uint64(tables.num_nodes), _p(tables.nodes_flags), _p(tables.nodes_time),
            C.c_uint64(tables.num_edges), _p(tables.edges_left), _p(tables.edges_right),
            _p(tables.edges_parent), _p(tables.edges_child),
            C.c_uint64(tables.num_sites), _p(tables.sites_position),
            _p(tables.sites_ancestral_state), _p(tables.sites_ancestral_state_offset),
            C.c_uint64(tables.num_mutations), _p(tables.mutations_site),
            _p(tables.mutations_node), _p(mp), _p(tables.mutations_derived_state),
            _p(tables.mutations_derived_state_offset)))
        self.h = h
        self.num_samples = int(L.ref_num_samples(h))
        self.num_trees = int(L.ref_num_trees(h))

    def __del__(self):
        if getattr(self, "h", None) is not None and _lib is not None:
            _lib.ref_treeseq_free(self.h)
            self.h = None

    def samples(self):
        out = np.empty(self.num_samples, dtype=np.int32)
        lib().ref_get_samples(self.h, _p(out))
        return out

    def breakpoints(self):
        out = np.empty(self.num_trees + 1, dtype=np.float64)
        lib().ref_get_breakpoints(self.h, _p(out))
        return out

    def indexes(self):
        E = self.t.num_edges
        i = np.empty(E, dtype=np.int32)
        o = np.empty(E, dtype=np.int32)
        lib().ref_get_indexes(self.h, _p(i), _p(o))
        return i, o

    def mutation_parents(self):
        out = np.empty(self.t.num_mutations, dtype=np.int32)
        lib().ref_get_mutation_parents(self.h, _p(out))
        return out

    @staticmethod
    def _sets(sample_sets):
        sizes = np.array([len(s) for s in sample_sets], dtype=np.uint64)
        flat = (np.concatenate([np.asarray(s, dtype=np.int32) for s in sample_sets])
                if len(sample_sets) else np.zeros(0, dtype=np.int32))
        return sizes, np.ascontiguousarray(flat, dtype=np.int32)

    def _windows(self, windows):
        if windows is None:
            windows = [0.0, self.t.sequence_length]
        return np.ascontiguousarray(windows, dtype=np.float64)

    def one_way(self, name, sample_sets, windows=None, mode="site", span_normalise=True,
                polarised=False):
        sizes, flat = self._sets(sample_sets)
        w = self._windows(windows)
        res = np.empty((len(w) - 1, len(sizes)), dtype=np.float64)
        _check(lib().ref_one_way_stat(
            self.h, C.c_int(ONE_WAY[name]), C.c_uint64(len(sizes)), _p(sizes), _p(flat),
            C.c_uint64(len(w) - 1), _p(w),
            C.c_uint32(mode_flags(mode, span_normalise, polarised)), _p(res)))
        return res

    def k_way(self, name, sample_sets, indexes, windows=None, mode="site",
              span_normalise=True, polarised=False, centre=True):
        sizes, flat = self._sets(sample_sets)
        w = self._windows(windows)
        idx = np.ascontiguousarray(indexes, dtype=np.int32).reshape(-1, TUPLE[name])
        res = np.empty((len(w) - 1, len(idx)), dtype=np.float64)
        _check(lib().ref_k_way_stat(
            self.h, C.c_int(K_WAY[name]), C.c_uint64(len(sizes)), _p(sizes), _p(flat),
            C.c_uint64(len(idx)), _p(idx), C.c_uint64(len(w) - 1), _p(w),
            C.c_uint32(mode_flags(mode, span_normalise, polarised, centre)), _p(res)))
        return res

    def general_stat(self, weights, f, output_dim, windows=None, mode="site",
                     span_normalise=True, polarised=False):
        """f: python callable (K,) -> (M,), called through a C trampoline exactly
        as _tskitmodule.c:6532-6586 does."""
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        K = weights.shape[1]
        w = self._windows(windows)
        res = np.empty((len(w) - 1, output_dim), dtype=np.float64)

        def tramp(k, state, m, result, params):
            x = np.ctypeslib.as_array(state, shape=(k,))
            y = np.asarray(f(x.copy()), dtype=np.float64)
            for j in range(m):
                result[j] = y[j]
            return 0

        cb = GENERAL_STAT_FUNC(tramp)
        _check(lib().ref_general_stat(
            self.h, C.c_uint64(K), _p(weights), C.c_uint64(output_dim), cb, None,
            C.c_uint64(len(w) - 1), _p(w),
            C.c_uint32(mode_flags(mode, span_normalise, polarised)), _p(res)))
        return res

    def divergence_matrix(self, sample_sets=None, windows=None, mode="site",
                          span_normalise=True):
        w = self._windows(windows)
        if sample_sets is None:
            sizes, flat, n = None, None, self.num_samples
        else:
            sizes, flat = self._sets(sample_sets)
            n = len(sizes)
        res = np.empty((len(w) - 1, n, n), dtype=np.float64)
        _check(lib().ref_divergence_matrix(
            self.h, C.c_uint64(n), _p(sizes), _p(flat), C.c_uint64(len(w) - 1), _p(w),
            C.c_uint32(mode_flags(mode, span_normalise)), _p(res)))
        return res

    def genotype_matrix(self, samples=None, isolated_as_missing=True):
        n = self.num_samples if samples is None else len(samples)
        if samples is not None:
            samples = np.ascontiguousarray(samples, dtype=np.int32)
        out = np.empty((self.t.num_sites, n), dtype=np.int32)
        _check(lib().ref_genotype_matrix(
            self.h, _p(samples), C.c_uint64(0 if samples is None else n),
            C.c_uint32(0 if isolated_as_missing else ISOLATED_NOT_MISSING), _p(out)))
        return out

    def trees_at(self, positions, tracked=None):
        """(parent[q, N], count[q, N]) of the tree covering each position."""
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        N = self.t.num_nodes
        par = np.empty((len(pos), N), dtype=np.int32)
        cnt = np.empty((len(pos), N), dtype=np.int32)
        if tracked is not None:
            tracked = np.ascontiguousarray(tracked, dtype=np.int32)
        _check(lib().ref_trees_at(
            self.h, C.c_uint64(len(pos)), _p(pos), _p(tracked),
            C.c_uint64(0 if tracked is None else len(tracked)), _p(par), _p(cnt)))
        return par, cnt


def sort_simplify(sequence_length, node_flags, node_time, left, right, parent, child,
                  samples):
    """tables.sort(); tables.simplify(samples) through the reference C library."""
    nf = np.ascontiguousarray(node_flags, dtype=np.uint32).copy()
    nt = np.ascontiguousarray(node_time, dtype=np.float64).copy()
    el = np.ascontiguousarray(left, dtype=np.float64).copy()
    er = np.ascontiguousarray(right, dtype=np.float64).copy()
    ep = np.ascontiguousarray(parent, dtype=np.int32).copy()
    ec = np.ascontiguousarray(child, dtype=np.int32).copy()
    s = np.ascontiguousarray(samples, dtype=np.int32)
    nn = C.c_uint64(len(nt))
    ne = C.c_uint64(len(el))
    node_map = np.empty(len(nt), dtype=np.int32)
    _check(lib().ref_sort_simplify(
        C.c_double(sequence_length), C.byref(nn), _p(nf), _p(nt), C.byref(ne), _p(el),
        _p(er), _p(ep), _p(ec), _p(s), C.c_uint64(len(s)), _p(node_map)))
    N, E = nn.value, ne.value
    return nf[:N], nt[:N], el[:E], er[:E], ep[:E], ec[:E], node_map
