"""Host-side mirror of the reference's low-level statistics interface.

``LLTreeSequence`` exposes the statistics methods of ``_tskit.TreeSequence``
(``python/_tskitmodule.c:6525-7510``; method table 8657-9022) with the same
names, positional/keyword arguments, dtypes, return shapes and exceptions, but
computes on the B200 through the C ABI in ``include/tskit_b200.h``.  It is what
the drop-in proxy (``tskit_b200.dropin``) puts in place of
``ts._ll_tree_sequence`` while a statistic runs.
"""
import ctypes as C
import threading

import numpy as np

from . import _lib
from .tables import Tables

STAT_SITE, STAT_BRANCH, STAT_NODE = 1, 2, 4
STAT_POLARISED, STAT_SPAN_NORMALISE = 1 << 10, 1 << 11
STAT_NONCENTRED = 1 << 14
ISOLATED_NOT_MISSING = 1 << 1
INIT_NODE_MODE = 1 << 0

STAT_IDS = {"diversity": 0, "segregating_sites": 1, "Y1": 2, "divergence": 3, "Y2": 4,
            "f2": 5, "genetic_relatedness": 6, "Y3": 7, "f3": 8, "f4": 9}
TUPLE = {"divergence": 2, "Y2": 2, "f2": 2, "genetic_relatedness": 2, "Y3": 3, "f3": 3,
         "f4": 4}


class LibraryError(Exception):
    """Mirror of ``_tskit.LibraryError`` (``_tskitmodule.c:231-300``)."""

    def __init__(self, code, msg=None):
        self.code = code
        super().__init__(msg if msg is not None else _lib.lib().tskb_strerror(code).decode())


def _handle(ret):
    if ret != 0:
        msg = _lib.lib().tskb_strerror(ret).decode()
        if ret == -20001:
            msg += ": " + _lib.lib().tskb_last_cuda_error().decode()
        raise LibraryError(ret, msg)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def parse_stats_mode(mode):
    """``parse_stats_mode`` (``_tskitmodule.c:6588-6607``)."""
    if mode is None or mode == "site":
        return STAT_SITE
    if mode == "branch":
        return STAT_BRANCH
    if mode == "node":
        return STAT_NODE
    if not isinstance(mode, str):
        raise TypeError("mode must be a string")
    raise ValueError("Unrecognised stats mode")


def parse_windows(windows):
    """``parse_windows`` (``_tskitmodule.c:6609-6636``)."""
    w = np.array(windows, dtype=np.float64, copy=True, order="C", ndmin=1)
    if w.ndim != 1:
        raise ValueError("object too deep for desired array")
    if w.shape[0] < 2:
        raise ValueError("Windows array must have at least 2 elements")
    return w


def parse_sample_sets(sample_set_sizes, sample_sets):
    """``parse_sample_sets`` (``_tskitmodule.c:799-851``)."""
    sizes = np.array(sample_set_sizes, dtype=np.uint64, copy=True, order="C", ndmin=1)
    sets = np.array(sample_sets, dtype=np.int32, copy=True, order="C", ndmin=1)
    if sizes.ndim != 1 or sets.ndim != 1:
        raise ValueError("object too deep for desired array")
    if int(sizes.sum()) != sets.shape[0]:
        raise ValueError("Sum of sample_set_sizes must equal length of sample_sets array")
    return sizes, sets


class LLTreeSequence:
    """Device-resident tree sequence exposing ``_tskit.TreeSequence``'s statistics methods."""

    def __init__(self, tables: Tables, device=0, genome_range=None, node_mode=False):
        tables.ensure_derived()
        self.tables = tables
        self.device = device
        # mode="node" statistics need a plan that keeps the pieces of parentless nodes too
        # (TSKB_INIT_NODE_MODE); it is staged on first use, next to the faster default plan
        self.node_mode = bool(node_mode)
        self._node_engine = None
        self._node_lock = threading.Lock()  # one node-mode engine, however many threads ask first
        lo, hi = (0.0, tables.sequence_length) if genome_range is None else genome_range
        self.genome_range = (float(lo), float(hi))
        t = _lib.Tables()
        t.sequence_length = tables.sequence_length
        t.time_uncalibrated = int(tables.time_uncalibrated)
        t.num_nodes = tables.num_nodes
        t.node_flags = _p(tables.nodes_flags)
        t.node_time = _p(tables.nodes_time)
        t.num_edges = tables.num_edges
        t.edge_left = _p(tables.edges_left)
        t.edge_right = _p(tables.edges_right)
        t.edge_parent = _p(tables.edges_parent)
        t.edge_child = _p(tables.edges_child)
        t.edge_insertion_order = _p(tables.edge_insertion_order)
        t.edge_removal_order = _p(tables.edge_removal_order)
        t.num_sites = tables.num_sites
        t.site_position = _p(tables.sites_position)
        t.site_ancestral_state = _p(tables.sites_ancestral_state)
        t.site_ancestral_state_offset = _p(tables.sites_ancestral_state_offset)
        t.num_mutations = tables.num_mutations
        t.mutation_site = _p(tables.mutations_site)
        t.mutation_node = _p(tables.mutations_node)
        t.mutation_parent = _p(tables.mutations_parent)
        t.mutation_derived_state = _p(tables.mutations_derived_state)
        t.mutation_derived_state_offset = _p(tables.mutations_derived_state_offset)
        h = C.c_void_p()
        self._h = None
        _handle(_lib.lib().tskb_treeseq_init(C.byref(h), C.byref(t), int(device),
                                             self.genome_range[0], self.genome_range[1],
                                             INIT_NODE_MODE if self.node_mode else 0))
        self._h = h

    def _for_mode(self, options):
        """The engine a call with these option bits runs on."""
        if (options & STAT_NODE) and not self.node_mode:
            with self._node_lock:
                if self._node_engine is None:
                    self._node_engine = LLTreeSequence(self.tables, device=self.device,
                                                       genome_range=self.genome_range, node_mode=True)
            return self._node_engine
        return self

    def close(self):
        if getattr(self, "_node_engine", None) is not None:
            self._node_engine.close()
            self._node_engine = None
        if getattr(self, "_h", None) is not None:
            _lib.lib().tskb_treeseq_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the accessors trees.py uses next to the stats methods
    def get_num_samples(self):
        return self.tables.num_samples

    def get_num_nodes(self):
        return self.tables.num_nodes

    def get_sequence_length(self):
        return self.tables.sequence_length

    def get_samples(self):
        return self.tables.samples

    def engine_stats(self):
        s = _lib.Stats()
        _handle(_lib.lib().tskb_treeseq_get_stats(self._h, C.byref(s)))
        return {
            "num_events": s.num_events, "num_visits": s.num_visits,
            "num_levels": s.num_levels, "stage_ms": s.stage_ms,
            "last_call_ms": s.last_call_ms, "last_kernel_ms": list(s.last_kernel_ms),
            "last_launches": s.last_launches, "device_bytes": s.device_bytes}

    def stat_device(self, name, sizes, d_sets_ptr, indexes, windows, options, d_result_ptr):
        """``tskb_treeseq_stat_device``: a sample-count statistic with the sample sets (int32) and the
        ``(W, M)`` result (float64) already resident in HBM, given as device addresses.  Synchronous:
        the result is complete when the call returns."""
        idx = None if indexes is None else np.ascontiguousarray(indexes, dtype=np.int32)
        sizes = np.ascontiguousarray(sizes, dtype=np.uint64)
        w = np.ascontiguousarray(windows, dtype=np.float64)
        _handle(_lib.lib().tskb_treeseq_stat_device(
            self._h, STAT_IDS[name], len(sizes), _p(sizes), C.c_void_p(d_sets_ptr),
            0 if idx is None else idx.shape[0], _p(idx), len(w) - 1, _p(w), options,
            C.c_void_p(d_result_ptr)))

    def debug_array(self, name, dtype):
        n = _lib.lib().tskb_treeseq_debug_array(self._h, name.encode(), None, 0)
        if n < 0:
            raise LibraryError(int(n))
        out = np.empty(n, dtype=dtype)
        _lib.lib().tskb_treeseq_debug_array(self._h, name.encode(), _p(out), out.nbytes)
        return out

    # ---- TreeSequence_one_way_stat_method (_tskitmodule.c:6909-6975)
    def _one_way(self, name, sample_set_sizes, sample_sets, windows, mode, span_normalise,
                 polarised):
        options = parse_stats_mode(mode)
        sizes, sets = parse_sample_sets(sample_set_sizes, sample_sets)
        w = parse_windows(windows)
        if span_normalise:
            options |= STAT_SPAN_NORMALISE
        if polarised:
            options |= STAT_POLARISED
        if options & STAT_NODE:
            result = np.zeros((len(w) - 1, self.tables.num_nodes, len(sizes)))
        else:
            result = np.zeros((len(w) - 1, len(sizes)))
        fn = getattr(_lib.lib(), "tskb_treeseq_" + name)
        _handle(fn(self._for_mode(options)._h, len(sizes), _p(sizes), _p(sets), len(w) - 1, _p(w), options,
                   _p(result)))
        return result

    def diversity(self, sample_set_sizes, sample_sets, windows=None, mode=None,
                  span_normalise=True, polarised=False):
        return self._one_way("diversity", sample_set_sizes, sample_sets, windows, mode,
                             span_normalise, polarised)

    def segregating_sites(self, sample_set_sizes, sample_sets, windows=None, mode=None,
                          span_normalise=True, polarised=False):
        return self._one_way("segregating_sites", sample_set_sizes, sample_sets, windows, mode,
                             span_normalise, polarised)

    def Y1(self, sample_set_sizes, sample_sets, windows=None, mode=None, span_normalise=True,
           polarised=False):
        return self._one_way("Y1", sample_set_sizes, sample_sets, windows, mode,
                             span_normalise, polarised)

    # ---- TreeSequence_k_way_stat_method (_tskitmodule.c:7104-7194)
    def _k_way(self, name, sample_set_sizes, sample_sets, indexes, windows, mode,
               span_normalise, polarised, centre):
        options = parse_stats_mode(mode)
        sizes, sets = parse_sample_sets(sample_set_sizes, sample_sets)
        w = parse_windows(windows)
        if span_normalise:
            options |= STAT_SPAN_NORMALISE
        if polarised:
            options |= STAT_POLARISED
        if not centre:
            options |= STAT_NONCENTRED
        idx = np.array(indexes, dtype=np.int32, copy=True, order="C")
        if idx.ndim != 2:
            raise ValueError("object of too small depth for desired array"
                             if idx.ndim < 2 else "object too deep for desired array")
        if idx.shape[0] < 1 or idx.shape[1] != TUPLE[name]:
            raise ValueError(f"indexes must be a k x {TUPLE[name]} array.")
        if options & STAT_NODE:
            result = np.zeros((len(w) - 1, self.tables.num_nodes, idx.shape[0]))
        else:
            result = np.zeros((len(w) - 1, idx.shape[0]))
        fn = getattr(_lib.lib(), "tskb_treeseq_" + name)
        _handle(fn(self._for_mode(options)._h, len(sizes), _p(sizes), _p(sets), idx.shape[0], _p(idx),
                   len(w) - 1, _p(w), options, _p(result)))
        return result

    def divergence(self, sample_set_sizes, sample_sets, indexes, windows=None, mode=None,
                   span_normalise=True, polarised=False, centre=True):
        return self._k_way("divergence", sample_set_sizes, sample_sets, indexes, windows, mode,
                           span_normalise, polarised, centre)

    def genetic_relatedness(self, sample_set_sizes, sample_sets, indexes, windows=None,
                            mode=None, span_normalise=True, polarised=False, centre=True):
        return self._k_way("genetic_relatedness", sample_set_sizes, sample_sets, indexes,
                           windows, mode, span_normalise, polarised, centre)

    def Y2(self, sample_set_sizes, sample_sets, indexes, windows=None, mode=None,
           span_normalise=True, polarised=False, centre=True):
        return self._k_way("Y2", sample_set_sizes, sample_sets, indexes, windows, mode,
                           span_normalise, polarised, centre)

    def f2(self, sample_set_sizes, sample_sets, indexes, windows=None, mode=None,
           span_normalise=True, polarised=False, centre=True):
        return self._k_way("f2", sample_set_sizes, sample_sets, indexes, windows, mode,
                           span_normalise, polarised, centre)

    def Y3(self, sample_set_sizes, sample_sets, indexes, windows=None, mode=None,
           span_normalise=True, polarised=False, centre=True):
        return self._k_way("Y3", sample_set_sizes, sample_sets, indexes, windows, mode,
                           span_normalise, polarised, centre)

    def f3(self, sample_set_sizes, sample_sets, indexes, windows=None, mode=None,
           span_normalise=True, polarised=False, centre=True):
        return self._k_way("f3", sample_set_sizes, sample_sets, indexes, windows, mode,
                           span_normalise, polarised, centre)

    def f4(self, sample_set_sizes, sample_sets, indexes, windows=None, mode=None,
           span_normalise=True, polarised=False, centre=True):
        return self._k_way("f4", sample_set_sizes, sample_sets, indexes, windows, mode,
                           span_normalise, polarised, centre)

    # ---- TreeSequence_general_stat (_tskitmodule.c:6665-6745)
    def general_stat(self, weights, summary_func, output_dim, windows=None, mode=None,
                     polarised=False, span_normalise=True):
        """0/1 weight columns only (what ``sample_count_stat`` builds,
        ``trees.py:8096``): one column is evaluated through a device lookup
        table of ``summary_func`` over every possible count.  Anything else
        raises; it is never computed on the host."""
        options = parse_stats_mode(mode)
        w = parse_windows(windows)
        if not callable(summary_func):
            raise TypeError("summary_func must be callable")
        W = np.array(weights, dtype=np.float64, copy=True, order="C")
        if W.ndim != 2:
            raise ValueError("object of too small depth for desired array")
        if W.shape[0] != self.tables.num_samples:
            raise ValueError("First dimension must be num_samples")
        if span_normalise:
            options |= STAT_SPAN_NORMALISE
        if polarised:
            options |= STAT_POLARISED
        if W.shape[1] != 1 or not np.all((W == 0) | (W == 1)):
            return self._general_stat_callback(W, summary_func, output_dim, w, options)
        samples = self.tables.samples
        members = samples[W[:, 0] == 1].astype(np.int32)
        sizes = np.array([len(members)], dtype=np.uint64)
        n = len(members)
        if n == 0:
            # the reference accepts an all-zero weight column (every state is 0); as a sample set it
            # would be TSK_ERR_EMPTY_SAMPLE_SET, so this call is handed back to the reference
            raise LibraryError(-20003, "general_stat with an all-zero weight column is not accelerated")
        table = np.empty((n + 1, output_dim), dtype=np.float64)
        for c in range(n + 1):
            y = np.asarray(summary_func(np.array([float(c)])), dtype=np.float64)
            if y.shape != (output_dim,):
                raise ValueError("summary_func returned array of wrong dimension")
            table[c] = y
        if options & STAT_NODE:
            result = np.zeros((len(w) - 1, self.tables.num_nodes, output_dim))
        else:
            result = np.zeros((len(w) - 1, output_dim))
        args = (1, _p(sizes), _p(members), output_dim, n + 1, _p(table), len(w) - 1, _p(w), options, _p(result))
        ret = _lib.lib().tskb_treeseq_sample_count_stat_tabulated(self._for_mode(options)._h, *args)
        if ret == -20003 and (options & STAT_BRANCH) and not np.isfinite(table).all():
            # NaN / inf summary values reach the running sum through nodes without a branch above them
            # too (0 x NaN): the engine that keeps every piece reproduces that
            ret = _lib.lib().tskb_treeseq_sample_count_stat_tabulated(self._for_mode(STAT_NODE)._h, *args)
        _handle(ret)
        return result

    def _general_stat_callback(self, W, summary_func, output_dim, w, options):
        """Several state columns or arbitrary weights: ``tskb_treeseq_general_stat`` with the Python
        callable behind a C callback (``general_stat_func``, ``_tskitmodule.c:6532-6586``).  The engine
        calls it once per distinct state vector, on the host."""
        if options & STAT_NODE:
            raise LibraryError(-20003, "node-mode general_stat with a Python summary function is not accelerated")
        if W.shape[1] > 8:
            raise LibraryError(-20003, "general_stat is accelerated for at most 8 state columns")
        K, M = W.shape[1], int(output_dim)
        failure = []

        def trampoline(k, state, m, out, params):
            try:
                y = np.asarray(summary_func(np.ctypeslib.as_array(state, shape=(k,)).copy()), dtype=np.float64)
                if y.shape != (m,):
                    raise ValueError("summary_func returned array of wrong dimension")
                np.ctypeslib.as_array(out, shape=(m,))[:] = y
                return 0
            except BaseException as e:  # handed back after the call, as _tskitmodule.c:6528, 6731 does
                failure.append(e)
                return -100000
        cb = _lib.GENERAL_STAT_FUNC(trampoline)
        result = np.zeros((len(w) - 1, M))
        args = (K, _p(W), M, cb, None, len(w) - 1, _p(w), options, _p(result))
        # branch mode runs on the engine that keeps every piece: a summary that is NaN / inf at the state
        # of a node without a branch above it must reach the running sum as 0 x NaN, as in the reference
        engine = self._for_mode(STAT_NODE) if (options & STAT_BRANCH) else self
        ret = _lib.lib().tskb_treeseq_general_stat(engine._h, *args)
        if failure:
            raise failure[0]
        _handle(ret)
        return result

    # ---- TreeSequence_allele_frequency_spectrum (_tskitmodule.c:6977-7063)
    def allele_frequency_spectrum(self, sample_set_sizes, sample_sets, windows, time_windows,
                                  mode=None, span_normalise=True, polarised=False):
        options = parse_stats_mode(mode)
        if span_normalise:
            options |= STAT_SPAN_NORMALISE
        if polarised:
            options |= STAT_POLARISED
        sizes, sets = parse_sample_sets(sample_set_sizes, sample_sets)
        w = parse_windows(windows)
        tw = parse_windows(time_windows)
        result = np.zeros([len(w) - 1, len(tw) - 1] + [int(x) + 1 for x in sizes])
        args = (len(sizes), _p(sizes), _p(sets), len(w) - 1, _p(w), len(tw) - 1, _p(tw), options, _p(result))
        ret = _lib.lib().tskb_treeseq_allele_frequency_spectrum(self._h, *args)
        if ret == -20003 and (options & STAT_BRANCH) and not self.node_mode:
            # time windows other than [0, inf) split every branch by time: the engine that keeps the
            # node of every piece (staged on first use) does it
            ret = _lib.lib().tskb_treeseq_allele_frequency_spectrum(self._for_mode(STAT_NODE)._h, *args)
        _handle(ret)
        return result

    # ---- TreeSequence_one_way_weighted_method (_tskitmodule.c:6747-6830)
    def _parse_weights(self, weights):
        W = np.array(weights, dtype=np.float64, copy=True, order="C")
        if W.ndim != 2:
            raise ValueError("object of too small depth for desired array"
                             if W.ndim < 2 else "object too deep for desired array")
        if W.shape[0] != self.tables.num_samples:
            raise ValueError("First dimension must be num_samples")
        return W

    def _one_way_weighted(self, name, weights, windows, mode, polarised, span_normalise):
        options = parse_stats_mode(mode)
        if polarised:
            options |= STAT_POLARISED
        if span_normalise:
            options |= STAT_SPAN_NORMALISE
        w = parse_windows(windows)
        W = self._parse_weights(weights)
        if options & STAT_NODE:
            result = np.zeros((len(w) - 1, self.tables.num_nodes, W.shape[1]))
        else:
            result = np.zeros((len(w) - 1, W.shape[1]))
        fn = getattr(_lib.lib(), "tskb_treeseq_" + name)
        _handle(fn(self._for_mode(options)._h, W.shape[1], _p(W), len(w) - 1, _p(w), options, _p(result)))
        return result

    def trait_covariance(self, weights, windows, mode=None, polarised=False, span_normalise=False):
        return self._one_way_weighted("trait_covariance", weights, windows, mode, polarised,
                                      span_normalise)

    def trait_correlation(self, weights, windows, mode=None, polarised=False, span_normalise=False):
        return self._one_way_weighted("trait_correlation", weights, windows, mode, polarised,
                                      span_normalise)

    # ---- TreeSequence_one_way_covariates_method (_tskitmodule.c:6819-6905)
    def trait_linear_model(self, weights, covariates, windows, mode=None, polarised=False,
                           span_normalise=False):
        options = parse_stats_mode(mode)
        if polarised:
            options |= STAT_POLARISED
        if span_normalise:
            options |= STAT_SPAN_NORMALISE
        w = parse_windows(windows)
        W = self._parse_weights(weights)
        Z = self._parse_weights(covariates)
        if options & STAT_NODE:
            result = np.zeros((len(w) - 1, self.tables.num_nodes, W.shape[1]))
        else:
            result = np.zeros((len(w) - 1, W.shape[1]))
        _handle(_lib.lib().tskb_treeseq_trait_linear_model(
            self._for_mode(options)._h, W.shape[1], _p(W), Z.shape[1], _p(Z), len(w) - 1, _p(w), options,
            _p(result)))
        return result

    # ---- TreeSequence_k_way_weighted_stat_method (_tskitmodule.c:7280-7388)
    def genetic_relatedness_weighted(self, weights, indexes, windows, mode=None,
                                     span_normalise=True, polarised=False, centre=True):
        options = parse_stats_mode(mode)
        if span_normalise:
            options |= STAT_SPAN_NORMALISE
        if polarised:
            options |= STAT_POLARISED
        if not centre:
            options |= STAT_NONCENTRED
        w = parse_windows(windows)
        W = self._parse_weights(weights)
        idx = np.array(indexes, dtype=np.int32, copy=True, order="C")
        if idx.ndim != 2:
            raise ValueError("object of too small depth for desired array"
                             if idx.ndim < 2 else "object too deep for desired array")
        if idx.shape[0] < 1 or idx.shape[1] != 2:
            raise ValueError("indexes must be a k x 2 array.")
        if options & STAT_NODE:
            result = np.zeros((len(w) - 1, self.tables.num_nodes, idx.shape[0]))
        else:
            result = np.zeros((len(w) - 1, idx.shape[0]))
        _handle(_lib.lib().tskb_treeseq_genetic_relatedness_weighted(
            self._for_mode(options)._h, W.shape[1], _p(W), idx.shape[0], _p(idx), len(w) - 1, _p(w), _p(result),
            options))
        return result

    # ---- TreeSequence_weighted_stat_vector_method (_tskitmodule.c:7197-7282)
    def genetic_relatedness_vector(self, weights, windows, mode=None, span_normalise=True,
                                   centre=True, nodes=None):
        options = parse_stats_mode(mode)
        if span_normalise:
            options |= STAT_SPAN_NORMALISE
        if not centre:
            options |= STAT_NONCENTRED
        w = parse_windows(windows)
        W = self._parse_weights(weights)
        focal = np.array(nodes, dtype=np.int32, copy=True, order="C")
        if focal.ndim != 1:
            raise ValueError("object of too small depth for desired array"
                             if focal.ndim < 1 else "object too deep for desired array")
        result = np.zeros((len(w) - 1, focal.shape[0], W.shape[1]))
        # focal nodes that are not samples: the engine that keeps the node of every piece
        engine = self
        in_range = (focal >= 0) & (focal < self.tables.num_nodes)
        if in_range.all() and focal.size and not (options & (STAT_SITE | STAT_NODE)):
            if not ((self.tables.nodes_flags[focal] & 1) != 0).all():  # NODE_IS_SAMPLE
                engine = self._for_mode(STAT_NODE)
        _handle(_lib.lib().tskb_treeseq_genetic_relatedness_vector(
            engine._h, W.shape[1], _p(W), len(w) - 1, _p(w), focal.shape[0], _p(focal), _p(result),
            options))
        return result

    # ---- TreeSequence_divergence_matrix (_tskitmodule.c:7435-7509)
    def divergence_matrix(self, windows, sample_sets=None, sample_set_sizes=None, mode=None,
                          span_normalise=True):
        options = parse_stats_mode(mode)
        w = parse_windows(windows)
        if span_normalise:
            options |= STAT_SPAN_NORMALISE
        if sample_sets is None:
            if sample_set_sizes is not None:
                raise TypeError("Must specify both sample_sets and sample_set_sizes")
            sizes = sets = None
            n = self.tables.num_samples
        else:
            if sample_set_sizes is None:
                raise TypeError("Must specify both sample_sets and sample_set_sizes")
            sizes, sets = parse_sample_sets(sample_set_sizes, sample_sets)
            n = len(sizes)
        result = np.zeros((len(w) - 1, n, n))
        _handle(_lib.lib().tskb_treeseq_divergence_matrix(
            self._h, n, _p(sizes), _p(sets), len(w) - 1, _p(w), options, _p(result)))
        return result

    # ---- integer parity outputs
    def trees_at(self, positions, tracked=None):
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        N = self.tables.num_nodes
        par = np.empty((len(pos), N), dtype=np.int32)
        cnt = np.empty((len(pos), N), dtype=np.int32)
        tr = None if tracked is None else np.ascontiguousarray(tracked, dtype=np.int32)
        _handle(_lib.lib().tskb_treeseq_trees_at(
            self._h, len(pos), _p(pos), _p(tr), 0 if tr is None else len(tr), _p(par),
            _p(cnt)))
        return par, cnt

    def matrix_phase_ms(self):
        """Phase times of the last site-mode divergence_matrix call (CUDA events)."""
        k = self.engine_stats()["last_kernel_ms"]
        return {"decode": k[0], "gemm": k[1], "finish": k[2], "alleles": int(k[7])}

    def decode_sites(self, first_site, num_sites, samples=None, isolated_as_missing=True):
        """int8 genotypes ``[num_sites, n]`` of the sites ``[first_site, first_site + num_sites)``."""
        n = self.tables.num_samples if samples is None else len(samples)
        s = None if samples is None else np.ascontiguousarray(samples, dtype=np.int32)
        out = np.empty((int(num_sites), n), dtype=np.int8)
        _handle(_lib.lib().tskb_treeseq_decode_sites(
            self._h, int(first_site), int(num_sites), _p(s), 0 if s is None else n,
            0 if isolated_as_missing else ISOLATED_NOT_MISSING, _p(out)))
        return out

    def genotype_matrix(self, samples=None, isolated_as_missing=True):
        n = self.tables.num_samples if samples is None else len(samples)
        s = None if samples is None else np.ascontiguousarray(samples, dtype=np.int32)
        out = np.empty((self.tables.num_sites, n), dtype=np.int8)
        _handle(_lib.lib().tskb_treeseq_genotype_matrix(
            self._h, _p(s), 0 if s is None else n,
            0 if isolated_as_missing else ISOLATED_NOT_MISSING, _p(out)))
        return out
