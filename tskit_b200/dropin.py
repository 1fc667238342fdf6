"""Drop-in seam behind the unchanged ``tskit.TreeSequence`` statistics API.

Every statistic of ``python/tskit/trees.py`` reaches C through the bound methods of
``self._ll_tree_sequence`` looked up at call time (``trees.py:4157-4158``; call sites
``:7997, 8551, 8620, 8679/8699/8799, 8948, 9837, 10010-10011, 10088/10098,
10182/10226/10265, 10321/10375/10421``).  ``AccelTreeSequence`` is a ``tskit.TreeSequence``
subclass sharing the same low-level object whose ``_ll_tree_sequence`` attribute is a property:
while one of the public statistics methods runs it returns a proxy that sends the low-level
statistics calls to the B200 engine (``lowlevel.LLTreeSequence``); at any other time it returns
the real ``_tskit.TreeSequence`` (whose C constructors type-check their argument,
``trees.py:713``, ``genotypes.py:124-125``), so trees(), variants(), tables, pickling ... are
untouched.  All of ``trees.py`` (argument normalisation, dimension dropping, Fst / Tajima's D /
GRM post-processing) runs unmodified on top.

    import tskit
    from tskit_b200.dropin import accelerate
    ts = accelerate(tskit.load("x.trees"))       # device staging happens once, here
    ts.diversity(sample_sets, windows=w, mode="branch")   # same call, same result, on the GPU

There is no CPU fallback for the accelerated calls: an engine failure raises.  Calls the engine
does not cover (float-weighted ``general_stat`` with a Python summary, branch-mode allele frequency
spectra with time windows other than ``[0, inf)``, more sample sets or output than one call holds)
are forwarded to the reference object and counted in ``ts.accel_stats["forwarded"]`` so that tests
can assert which path ran.  ``genetic_relatedness_vector`` is routed too, so ``ts.pca`` -- whose
randomised SVD iterates that product (``trees.py:9284-9557``) -- runs its products on the device.
"""
import threading

import numpy as np

from . import lowlevel
from .tables import Tables

try:  # the host product; tests put baseline/_ref on sys.path
    import tskit
    import _tskit
except ImportError:  # pragma: no cover
    tskit = None
    _tskit = None

ONE_WAY = ("diversity", "segregating_sites", "Y1")
K_WAY = ("divergence", "genetic_relatedness", "Y2", "f2", "Y3", "f3", "f4")
PUBLIC_STATS = (
    "diversity", "divergence", "divergence_matrix", "genetic_relatedness",
    "genetic_relatedness_matrix", "segregating_sites", "Tajimas_D", "Fst", "Y1", "Y2", "Y3",
    "f2", "f3", "f4", "sample_count_stat", "general_stat", "trait_covariance", "trait_correlation",
    "trait_linear_model", "genetic_relatedness_weighted", "allele_frequency_spectrum",
    "genetic_relatedness_vector")


def tables_from_tree_sequence(ts):
    """The columns the engine reads, as zero-copy numpy views where tskit offers them
    (``trees.py:4180-4209``)."""
    t = ts.tables
    return Tables(
        sequence_length=float(ts.sequence_length),
        nodes_flags=ts.nodes_flags, nodes_time=ts.nodes_time,
        edges_left=ts.edges_left, edges_right=ts.edges_right,
        edges_parent=ts.edges_parent, edges_child=ts.edges_child,
        sites_position=ts.sites_position,
        sites_ancestral_state=t.sites.ancestral_state,
        sites_ancestral_state_offset=t.sites.ancestral_state_offset,
        mutations_site=ts.mutations_site, mutations_node=ts.mutations_node,
        mutations_parent=ts.mutations_parent,
        mutations_derived_state=t.mutations.derived_state,
        mutations_derived_state_offset=t.mutations.derived_state_offset,
        time_uncalibrated=(ts.time_units == "uncalibrated"),
        edge_insertion_order=ts.indexes_edge_insertion_order,
        edge_removal_order=ts.indexes_edge_removal_order)


class _Proxy:
    """Stands in for ``ts._ll_tree_sequence`` while a statistics method runs."""

    def __init__(self, real, engine, counters):
        self._real = real
        self._engine = engine
        self._counters = counters

    def __getattr__(self, name):  # everything that is not a statistic
        return getattr(self._real, name)

    def _run(self, name, args, kwargs):
        try:
            out = getattr(self._engine, name)(*args, **kwargs)
        except lowlevel.LibraryError as e:
            if e.code == -20003:  # valid tskit call the engine does not accelerate
                self._counters["forwarded"] += 1
                return getattr(self._real, name)(*args, **kwargs)
            if -20000 < e.code < 0 and _tskit is not None:
                raise _tskit.LibraryError(str(e)) from None  # same type and text as the reference
            raise
        self._counters["accelerated"] += 1
        return out


def _make(name):
    def method(self, *args, **kwargs):
        return self._run(name, args, kwargs)
    method.__name__ = name  # trees.py:8195 inspects ll_method.__name__
    return method


WEIGHTED = ("trait_covariance", "trait_correlation", "trait_linear_model",
            "genetic_relatedness_weighted")

for _n in ONE_WAY + K_WAY + WEIGHTED + ("divergence_matrix", "general_stat", "allele_frequency_spectrum",
                                           "genetic_relatedness_vector"):
    setattr(_Proxy, _n, _make(_n))


if tskit is not None:

    class AccelTreeSequence(tskit.TreeSequence):
        """``tskit.TreeSequence`` whose statistics run on the B200 engine."""

        def __init__(self, ll_tree_sequence, device=0, engine=None):
            # Depth of nested public statistics calls, PER THREAD: only the thread inside a statistics
            # method sees the proxy; any other thread using the same object concurrently (ts.first(),
            # ts.variants(), tskit.Tree(ts): C constructors that type-check their argument) keeps
            # seeing the real _tskit.TreeSequence.
            self._accel_tls = threading.local()
            self._accel_proxy = None
            self.accel_stats = {"accelerated": 0, "forwarded": 0}
            super().__init__(ll_tree_sequence)
            if engine is None:
                engine = lowlevel.LLTreeSequence(tables_from_tree_sequence(self), device=device)
            self._accel_engine = engine
            self._accel_proxy = _Proxy(self._ll_real, engine, self.accel_stats)

        @property
        def _ll_tree_sequence(self):
            if self._accel_proxy is not None and getattr(self._accel_tls, "depth", 0) > 0:
                return self._accel_proxy
            return self._ll_real

        @_ll_tree_sequence.setter
        def _ll_tree_sequence(self, value):  # assigned by TreeSequence.__init__ (trees.py:4157)
            self._ll_real = value

        def _accel_call(self, name, args, kwargs):
            tls = self._accel_tls
            tls.depth = getattr(tls, "depth", 0) + 1
            try:
                if name in ("divergence_matrix", "genetic_relatedness_matrix") and kwargs.get("num_threads"):
                    # the reference splits the windows / trees over a thread pool whose workers look
                    # the low-level object up themselves (trees.py:8671-8706); one device call does
                    # the whole job, so the chunking is dropped ("Threading")
                    kwargs = dict(kwargs, num_threads=0)
                return getattr(super(), name)(*args, **kwargs)
            finally:
                tls.depth -= 1

        def __reduce__(self):
            return (_unpickle, (super().__reduce__(),))

    def _public(name):
        def method(self, *args, **kwargs):
            return self._accel_call(name, args, kwargs)
        method.__name__ = name
        method.__doc__ = getattr(tskit.TreeSequence, name).__doc__
        return method

    for _n in PUBLIC_STATS:
        setattr(AccelTreeSequence, _n, _public(_n))

    def _genotype_matrix(self, *, samples=None, isolated_as_missing=None, alleles=None,
                         impute_missing_data=None):
        """``TreeSequence.genotype_matrix`` (``trees.py:5563-5645``) from the device decode
        (``genotypes.c:473-594``) when the allele coding is the default one (ancestral state 0,
        derived states in mutation order); a user-supplied ``alleles`` mapping or the deprecated
        ``impute_missing_data`` go to the reference."""
        engine = self._accel_engine
        if (alleles is not None or impute_missing_data is not None
                or not isinstance(engine, lowlevel.LLTreeSequence)):
            self.accel_stats["forwarded"] += 1
            return tskit.TreeSequence.genotype_matrix(
                self, samples=samples, isolated_as_missing=isolated_as_missing, alleles=alleles,
                impute_missing_data=impute_missing_data)
        if samples is not None:
            samples = np.asarray(samples, dtype=np.int32)
        g = engine.genotype_matrix(samples=samples,
                                   isolated_as_missing=True if isolated_as_missing is None
                                   else bool(isolated_as_missing))
        self.accel_stats["accelerated"] += 1
        return g.astype(np.int32)
    _genotype_matrix.__name__ = "genotype_matrix"
    AccelTreeSequence.genotype_matrix = _genotype_matrix

    class _DeviceLLVariant:
        """Stands in for ``_tskit.Variant`` inside a ``tskit.Variant`` (``genotypes.py:118-127`` is the
        only place that constructs one): ``decode(site_id)`` serves the genotypes from blocks of sites
        decoded on the device (``tskb_treeseq_decode_sites``), the alleles from the tables in the
        reference's order -- ancestral state first, derived states in order of first appearance among
        the site's mutations, ``None`` appended when a requested node is missing
        (``genotypes.c:533-594``)."""

        BLOCK_BYTES = 256 << 20

        def __init__(self, acc, samples, isolated_as_missing):
            self._acc = acc
            self._engine = acc._accel_engine
            t = acc._accel_engine.tables
            self.samples = t.samples.copy() if samples is None else np.array(samples, dtype=np.int32)
            self._samples_arg = None if samples is None else self.samples
            self.isolated_as_missing = bool(isolated_as_missing)
            self.site_id = tskit.NULL
            self.genotypes = np.zeros(len(self.samples), dtype=np.int32)
            self.alleles = ()
            self._lo = self._hi = 0
            self._block = None
            n = max(1, len(self.samples))
            self._block_sites = max(1, min(t.num_sites, self.BLOCK_BYTES // n))
            # mutation rows of every site (mutation_site is sorted)
            self._moff = np.searchsorted(t.mutations_site, np.arange(t.num_sites + 1))

        def _allele_strings(self, site_id):
            t = self._engine.tables
            a, ao = t.sites_ancestral_state, t.sites_ancestral_state_offset
            d, do = t.mutations_derived_state, t.mutations_derived_state_offset
            out = [a[int(ao[site_id]):int(ao[site_id + 1])].tobytes().decode()]
            for m in range(self._moff[site_id], self._moff[site_id + 1]):
                x = d[int(do[m]):int(do[m + 1])].tobytes().decode()
                if x not in out:
                    out.append(x)
            return out

        def decode(self, site_id):
            t = self._engine.tables
            site_id = int(site_id)
            if site_id < 0 or site_id >= t.num_sites:
                raise _tskit.LibraryError("Site out of bounds. (TSK_ERR_SITE_OUT_OF_BOUNDS)")
            if self._block is None or not (self._lo <= site_id < self._hi):
                self._lo = site_id
                self._hi = min(t.num_sites, site_id + self._block_sites)
                self._block = self._engine.decode_sites(self._lo, self._hi - self._lo, samples=self._samples_arg,
                                                        isolated_as_missing=self.isolated_as_missing)
                self._acc.accel_stats["accelerated"] += 1
            self.genotypes = self._block[site_id - self._lo].astype(np.int32)
            alleles = self._allele_strings(site_id)
            if (self.genotypes == tskit.MISSING_DATA).any():
                alleles.append(None)
            self.alleles = tuple(alleles)
            self.site_id = site_id

        def restricted_copy(self):
            c = _DeviceLLVariant.__new__(_DeviceLLVariant)
            c.__dict__.update(self.__dict__)
            c.genotypes = self.genotypes.copy()
            c._block = None  # a copy keeps its site; decoding it again is an error in the reference too
            c.decode = c._no_decode
            return c

        def _no_decode(self, site_id):
            raise _tskit.LibraryError("Can't decode a copy of a variant. (TSK_ERR_VARIANT_CANT_DECODE_COPY)")

    def _variants(self, *, samples=None, isolated_as_missing=None, alleles=None, impute_missing_data=None,
                  copy=None, left=None, right=None):
        """``TreeSequence.variants`` (``trees.py:5444-5560``) over the device decode: the iterator yields
        real ``tskit.Variant`` objects whose low-level half is ``_DeviceLLVariant``; ``haplotypes`` and
        ``alignments`` iterate this method and so run on the device decode too.  A user-supplied
        ``alleles`` coding, the deprecated ``impute_missing_data`` and requested nodes that are not
        samples go to the reference."""
        engine = self._accel_engine
        forward = (alleles is not None or impute_missing_data is not None
                   or not isinstance(engine, lowlevel.LLTreeSequence))
        if samples is not None and not forward:
            sm = np.asarray(samples, dtype=np.int64)
            ok = sm.ndim == 1 and ((sm >= 0) & (sm < self.num_nodes)).all()
            forward = not ok or not ((self.nodes_flags[sm] & 1) != 0).all()
        if forward:
            self.accel_stats["forwarded"] += 1
            yield from tskit.TreeSequence.variants(
                self, samples=samples, isolated_as_missing=isolated_as_missing, alleles=alleles,
                impute_missing_data=impute_missing_data, copy=copy, left=left, right=right)
            return
        interval = self._check_genomic_range(left, right)
        if isolated_as_missing is None:
            isolated_as_missing = True
        if copy is None:
            copy = True
        variant = tskit.Variant.__new__(tskit.Variant)
        variant.tree_sequence = self
        variant._ll_variant = _DeviceLLVariant(self, samples, isolated_as_missing)
        start, stop = np.searchsorted(self.sites_position, interval)
        for site_id in range(int(start), int(stop)):
            variant.decode(site_id)
            yield variant.copy() if copy else variant
    _variants.__name__ = "variants"
    AccelTreeSequence.variants = _variants

    def _unpickle(base):
        fn, args = base[0], base[1]
        return accelerate(fn(*args))


def accelerate(ts, device=0, engine=None):
    """Wrap a ``tskit.TreeSequence`` (same low-level object, no copy of the tables on the host)."""
    if tskit is None:
        raise RuntimeError("tskit is not importable")
    return AccelTreeSequence(ts.ll_tree_sequence, device=device, engine=engine)


def from_tables(tables: Tables):
    """Build a reference ``tskit.TreeSequence`` from a ``Tables`` container (tests, benchmarks)."""
    tc = tskit.TableCollection(tables.sequence_length)
    if tables.time_uncalibrated:
        tc.time_units = "uncalibrated"
    N = tables.num_nodes
    tc.nodes.set_columns(flags=tables.nodes_flags, time=tables.nodes_time,
                         population=np.full(N, -1, dtype=np.int32),
                         individual=np.full(N, -1, dtype=np.int32))
    tc.edges.set_columns(left=tables.edges_left, right=tables.edges_right,
                         parent=tables.edges_parent, child=tables.edges_child)
    if tables.num_sites:
        tc.sites.set_columns(position=tables.sites_position,
                             ancestral_state=tables.sites_ancestral_state,
                             ancestral_state_offset=tables.sites_ancestral_state_offset)
        tables.ensure_derived()
        Mu = tables.num_mutations
        tc.mutations.set_columns(site=tables.mutations_site, node=tables.mutations_node,
                                 parent=tables.mutations_parent,
                                 time=np.full(Mu, tskit.UNKNOWN_TIME),
                                 derived_state=tables.mutations_derived_state,
                                 derived_state_offset=tables.mutations_derived_state_offset)
    tc.sort()
    return tc.tree_sequence()
