"""The drop-in seam (tskit_b200/dropin.py) under the unchanged tskit.TreeSequence API.
CPU part: seam mechanics with an engine stand-in that answers from the reference's own low-level
object (no CUDA).  GPU part: every public statistic through the real engine == the reference."""
import numpy as np
import pytest

tskit = pytest.importorskip("tskit", reason="baseline/_ref (reference tskit) not installed")

from tskit_b200 import dropin  # noqa: E402


class EchoEngine:
    """Engine stand-in: forwards to the reference low-level object and records the calls."""

    def __init__(self, ll):
        self.ll = ll
        self.calls = []

    def __getattr__(self, name):
        def f(*a, **k):
            self.calls.append(name)
            return getattr(self.ll, name)(*a, **k)
        f.__name__ = name
        return f


@pytest.fixture(scope="module")
def ts(wf_small):
    return dropin.from_tables(wf_small)


def test_seam_routes_statistics_only(ts):
    eng = EchoEngine(ts.ll_tree_sequence)
    acc = dropin.accelerate(ts, engine=eng)
    s = ts.samples()
    sets = [s[:60], s[60:130], s[130:]]
    w = np.linspace(0, ts.sequence_length, 6)
    assert np.array_equal(acc.diversity(sets, windows=w, mode="branch"),
                          ts.diversity(sets, windows=w, mode="branch"))
    assert np.array_equal(acc.Fst(sets, indexes=[(0, 1), (1, 2)], windows=w),
                          ts.Fst(sets, indexes=[(0, 1), (1, 2)], windows=w))
    assert np.array_equal(acc.Tajimas_D(sets, windows=w), ts.Tajimas_D(sets, windows=w))
    assert np.array_equal(acc.f4(sets + [s[:10]], indexes=[(0, 1, 2, 3)], mode="branch"),
                          ts.f4(sets + [s[:10]], indexes=[(0, 1, 2, 3)], mode="branch"))
    assert np.array_equal(acc.genetic_relatedness(sets, indexes=[(0, 1)], mode="site"),
                          ts.genetic_relatedness(sets, indexes=[(0, 1)], mode="site"))
    assert "diversity" in eng.calls and "divergence" in eng.calls and "f4" in eng.calls
    assert "segregating_sites" in eng.calls  # Tajima's D and relatedness(proportion=True)
    assert acc.accel_stats["accelerated"] == len(eng.calls)
    # outside a statistics call the real low-level object is visible: C constructors work
    assert acc._ll_tree_sequence is ts.ll_tree_sequence
    assert acc.first().num_samples() == ts.num_samples
    assert sum(1 for _ in acc.variants()) == ts.num_sites
    n_before = len(eng.calls)
    acc.genotype_matrix()
    acc.simplify(s[:20])
    assert len(eng.calls) == n_before


def test_node_mode_goes_to_the_engine(ts):
    eng = EchoEngine(ts.ll_tree_sequence)
    acc = dropin.accelerate(ts, engine=eng)
    got = acc.diversity([ts.samples()[:50]], mode="node")
    assert got.shape == (ts.num_nodes, 1)
    assert acc.accel_stats == {"accelerated": 1, "forwarded": 0}
    assert eng.calls == ["diversity"]


def test_seam_is_per_thread(ts):
    """While one thread is inside a statistic, other threads using the same object keep seeing the
    real low-level object (the C constructors of Tree / Variant type-check it)."""
    import threading
    import time

    inside, release = threading.Event(), threading.Event()

    class SlowEngine(EchoEngine):
        def __getattr__(self, name):
            f = EchoEngine.__getattr__(self, name)

            def g(*a, **k):
                inside.set()
                assert release.wait(30)
                return f(*a, **k)
            g.__name__ = name
            return g

    acc = dropin.accelerate(ts, engine=SlowEngine(ts.ll_tree_sequence))
    out = {}
    worker = threading.Thread(target=lambda: out.setdefault("pi", acc.diversity(mode="branch")))
    worker.start()
    assert inside.wait(30)
    try:
        # the statistic is in flight on the other thread: this thread must not see the proxy
        assert acc._ll_tree_sequence is ts.ll_tree_sequence
        assert acc.first().num_samples() == ts.num_samples
        assert sum(1 for _ in acc.variants()) == ts.num_sites
        assert tskit.Tree(acc).tree_sequence is acc
    finally:
        release.set()
        worker.join(30)
    assert np.array_equal(out["pi"], ts.diversity(mode="branch"))
    # the reference's thread-pool chunking (trees.py:8671-8706) is replaced by one engine call
    eng = EchoEngine(ts.ll_tree_sequence)
    acc2 = dropin.accelerate(ts, engine=eng)
    s = ts.samples()[:12]
    w = np.linspace(0, ts.sequence_length, 5)
    got = acc2.divergence_matrix([[int(u)] for u in s], windows=w, num_threads=3, mode="site")
    assert np.array_equal(got, ts.divergence_matrix([[int(u)] for u in s], windows=w, mode="site"))
    assert eng.calls == ["divergence_matrix"]
    time.sleep(0)


def test_errors_keep_reference_type(ts):
    import _tskit
    acc = dropin.accelerate(ts, engine=EchoEngine(ts.ll_tree_sequence))
    with pytest.raises(_tskit.LibraryError):
        acc.diversity([ts.samples()[:5]], windows=[0, ts.sequence_length / 2], mode="branch")


@pytest.mark.gpu
def test_public_api_matches_reference_on_gpu(ts):
    acc = dropin.accelerate(ts)
    s = ts.samples()
    sets = [s[:60], s[60:130], s[130:]]
    w = np.linspace(0, ts.sequence_length, 11)

    def same(a, b, tol=1e-9):
        a, b = np.asarray(a), np.asarray(b)
        assert a.shape == b.shape
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.allclose(a, b, rtol=tol, atol=tol * np.nanmax(np.abs(b), initial=0.0), equal_nan=True)

    for mode in ("site", "branch"):
        for windows in (None, w, "trees"):
            same(acc.diversity(sets, windows=windows, mode=mode), ts.diversity(sets, windows=windows, mode=mode))
            same(acc.divergence(sets, indexes=[(0, 1), (0, 2)], windows=windows, mode=mode),
                 ts.divergence(sets, indexes=[(0, 1), (0, 2)], windows=windows, mode=mode))
        same(acc.segregating_sites(sets, windows=w, mode=mode), ts.segregating_sites(sets, windows=w, mode=mode))
        same(acc.Fst(sets, indexes=[(0, 1), (1, 2)], windows=w, mode=mode),
             ts.Fst(sets, indexes=[(0, 1), (1, 2)], windows=w, mode=mode))
        same(acc.Tajimas_D(sets, windows=w, mode=mode), ts.Tajimas_D(sets, windows=w, mode=mode))
        for name, k in (("Y1", 1), ("Y2", 2), ("f2", 2), ("Y3", 3), ("f3", 3), ("f4", 4)):
            kw = {} if k == 1 else {"indexes": [tuple(range(k)) if k <= 3 else (0, 1, 2, 0)]}
            same(getattr(acc, name)(sets, windows=w, mode=mode, **kw),
                 getattr(ts, name)(sets, windows=w, mode=mode, **kw))
        same(acc.genetic_relatedness(sets, indexes=[(0, 1), (2, 2)], windows=w, mode=mode),
             ts.genetic_relatedness(sets, indexes=[(0, 1), (2, 2)], windows=w, mode=mode))
        same(acc.genetic_relatedness(sets, indexes=[(0, 1)], windows=w, mode=mode, proportion=False, centre=False),
             ts.genetic_relatedness(sets, indexes=[(0, 1)], windows=w, mode=mode, proportion=False, centre=False))
        n = len(sets[0])
        f = lambda x: x * (n - x) / (n * (n - 1))  # noqa: E731
        same(acc.sample_count_stat([sets[0]], f, 1, windows=w, mode=mode),
             ts.sample_count_stat([sets[0]], f, 1, windows=w, mode=mode))
    assert acc.accel_stats["forwarded"] == 0 and acc.accel_stats["accelerated"] > 30
    # node mode: every node has a value in every window (trees.c:1788-1918); a second plan that keeps
    # the pieces of parentless nodes is staged on first use
    same(acc.diversity([s[:50], s[50:]], windows=w, mode="node"),
         ts.diversity([s[:50], s[50:]], windows=w, mode="node"))
    same(acc.divergence(sets, indexes=[(0, 1), (2, 0)], windows=w, mode="node", span_normalise=False),
         ts.divergence(sets, indexes=[(0, 1), (2, 0)], windows=w, mode="node", span_normalise=False))
    same(acc.f3(sets, indexes=[(0, 1, 2)], mode="node"), ts.f3(sets, indexes=[(0, 1, 2)], mode="node"))
    same(acc.Y1(sets, windows=w, mode="node"), ts.Y1(sets, windows=w, mode="node"))
    W2 = np.random.default_rng(3).normal(size=(ts.num_samples, 2))
    same(acc.trait_covariance(W2, windows=w, mode="node"), ts.trait_covariance(W2, windows=w, mode="node"))
    assert acc.accel_stats["forwarded"] == 0
    # genotype decode behind the unchanged TreeSequence.genotype_matrix
    assert np.array_equal(acc.genotype_matrix(), ts.genotype_matrix())
    sub = s[[7, 3, 150, 20]]
    assert np.array_equal(acc.genotype_matrix(samples=sub, isolated_as_missing=False),
                          ts.genotype_matrix(samples=sub, isolated_as_missing=False))
    assert acc.genotype_matrix().dtype == ts.genotype_matrix().dtype
    assert acc.accel_stats["forwarded"] == 0
    assert np.array_equal(acc.genotype_matrix(alleles=("0", "1")), ts.genotype_matrix(alleles=("0", "1")))
    assert acc.accel_stats["forwarded"] == 1
    assert acc.first().num_samples() == ts.num_samples


def test_oracle_weighted_statistics_pinned_to_reference(ts, wf_small):
    """The oracle's weighted statistics (oracle/port.py, numpy restatement of trees.c:3960-4110,
    4800-4897 on top of its general_stat) against the reference package itself (CPU)."""
    from oracle import port
    o = port.Oracle(wf_small)
    rng = np.random.default_rng(5)
    W = rng.normal(size=(ts.num_samples, 3)) + np.array([0.0, 2.0, -1.0])
    w = np.linspace(0, ts.sequence_length, 5)
    for mode in ("site", "branch"):
        assert np.allclose(o.trait_covariance(W, windows=w, mode=mode),
                           ts.trait_covariance(W, windows=w, mode=mode), rtol=1e-9, atol=1e-12)
        assert np.allclose(o.trait_correlation(W, windows=w, mode=mode),
                           ts.trait_correlation(W, windows=w, mode=mode), rtol=1e-9, atol=1e-12)
        Z = np.linalg.qr(np.column_stack([np.ones(ts.num_samples), rng.normal(size=(ts.num_samples, 2))]))[0]
        got = o.trait_linear_model(W, Z, windows=w, mode=mode)
        want = ts.ll_tree_sequence.trait_linear_model(W, Z, w, mode=mode, span_normalise=True)
        assert np.allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())
        idx = [(0, 1), (2, 2), (1, 0)]
        for centre in (True, False):
            got = o.genetic_relatedness_weighted(W, idx, windows=w, mode=mode, centre=centre)
            want = ts.genetic_relatedness_weighted(W, indexes=idx, windows=w, mode=mode, centre=centre)
            assert np.allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max()), (mode, centre)


@pytest.mark.gpu
def test_weighted_statistics_through_dropin(ts):
    acc = dropin.accelerate(ts)
    rng = np.random.default_rng(7)
    W = rng.normal(size=(ts.num_samples, 2)) * np.array([1.0, 30.0]) + np.array([5.0, -2.0])
    w = np.linspace(0, ts.sequence_length, 7)

    def same(a, b):
        assert np.allclose(a, b, rtol=1e-9, atol=1e-9 * np.abs(b).max())
    for mode in ("site", "branch"):
        same(acc.trait_covariance(W, windows=w, mode=mode), ts.trait_covariance(W, windows=w, mode=mode))
        same(acc.trait_correlation(W, windows=w, mode=mode), ts.trait_correlation(W, windows=w, mode=mode))
        same(acc.trait_covariance(W[:, :1], mode=mode, span_normalise=False),
             ts.trait_covariance(W[:, :1], mode=mode, span_normalise=False))
        for centre in (True, False):
            same(acc.genetic_relatedness_weighted(W, indexes=[(0, 1), (1, 1)], windows=w, mode=mode, centre=centre),
                 ts.genetic_relatedness_weighted(W, indexes=[(0, 1), (1, 1)], windows=w, mode=mode, centre=centre))
        Z = rng.normal(size=(ts.num_samples, 2))
        same(acc.trait_linear_model(W, Z, windows=w, mode=mode), ts.trait_linear_model(W, Z, windows=w, mode=mode))
        same(acc.trait_linear_model(W[:, :1], None, mode=mode), ts.trait_linear_model(W[:, :1], None, mode=mode))
    assert acc.accel_stats["forwarded"] == 0 and acc.accel_stats["accelerated"] == 14
    W9 = rng.normal(size=(ts.num_samples, 9))  # more state columns than a sweep carries: batched
    same(acc.trait_covariance(W9, mode="branch"), ts.trait_covariance(W9, mode="branch"))
    same(acc.trait_correlation(W9, mode="site"), ts.trait_correlation(W9, mode="site"))
    pairs = [(0, 8), (3, 3), (8, 1), (5, 6), (2, 7), (4, 0)]
    same(acc.genetic_relatedness_weighted(W9, indexes=pairs, mode="branch"),
         ts.genetic_relatedness_weighted(W9, indexes=pairs, mode="branch"))
    assert acc.accel_stats["forwarded"] == 0


def test_oracle_site_afs_pinned_to_reference(ts, wf_small):
    """The oracle's by-definition site-mode AFS against the reference package (CPU)."""
    from oracle import port
    o = port.Oracle(wf_small)
    s = ts.samples()
    sets = [s[:3], s[3:7]]
    w = np.linspace(0, ts.sequence_length, 4)
    for pol in (False, True):
        for span in (True, False):
            got = o.site_allele_frequency_spectrum(sets, windows=w, polarised=pol, span_normalise=span)
            want = ts.allele_frequency_spectrum(sets, windows=w, mode="site", polarised=pol, span_normalise=span)
            assert got.shape == want.shape and np.array_equal(got, want), (pol, span)
    assert np.array_equal(o.site_allele_frequency_spectrum([s], polarised=True)[0],
                          ts.allele_frequency_spectrum([s], mode="site", polarised=True))


def test_oracle_branch_afs_pinned_to_reference(ts, wf_small):
    """The oracle's step-by-step restatement of the branch-mode AFS, `last_update` bookkeeping
    included (several roots: nodes regain parents), against the reference package (CPU)."""
    from oracle import port
    from tests import fixtures as fx
    o = port.Oracle(wf_small)
    s = ts.samples()
    sets = [s[:3], s[3:7]]
    w = np.linspace(0, ts.sequence_length, 4)
    for pol in (False, True):
        for span in (True, False):
            got = o.branch_allele_frequency_spectrum(sets, windows=w, polarised=pol, span_normalise=span)
            want = ts.allele_frequency_spectrum(sets, windows=w, mode="branch", polarised=pol, span_normalise=span)
            assert got.shape == want.shape
            assert np.allclose(got, want, rtol=1e-12, atol=1e-12 * np.abs(want).max()), (pol, span)
    t = fx.load("multiroot")
    mts = dropin.from_tables(t)
    ms = mts.samples()
    got = port.Oracle(t).branch_allele_frequency_spectrum([ms[:2], ms[2:]], polarised=True, span_normalise=False)
    want = mts.allele_frequency_spectrum([ms[:2], ms[2:]], mode="branch", polarised=True, span_normalise=False)
    assert np.allclose(got[0], want, rtol=1e-12)


@pytest.mark.gpu
def test_site_afs_through_dropin(ts, wf_small):
    from oracle import port
    acc = dropin.accelerate(ts)
    o = port.Oracle(wf_small)
    s = ts.samples()
    w = np.linspace(0, ts.sequence_length, 6)
    for sets in ([s], [s[:50], s[50:]], [s[:3], s[3:10], s[20:22]], [s[::2]]):
        for pol in (False, True):
            for span in (True, False):
                got = acc.allele_frequency_spectrum(sets, windows=w, mode="site", polarised=pol, span_normalise=span)
                want = ts.allele_frequency_spectrum(sets, windows=w, mode="site", polarised=pol, span_normalise=span)
                assert got.shape == want.shape and np.allclose(got, want, rtol=1e-12, atol=0), (len(sets), pol, span)
    got = acc.allele_frequency_spectrum([s[:4], s[4:9]], mode="site", polarised=True, span_normalise=False)
    assert np.array_equal(got, o.site_allele_frequency_spectrum([s[:4], s[4:9]], polarised=True, span_normalise=False)[0])
    # branch mode, on an input with several roots (nodes regain parents: the reference credits them
    # from their last update, trees.c:3650-3697)
    for sets in ([s], [s[:50], s[50:]], [s[:3], s[3:10], s[20:22]]):
        for pol in (False, True):
            for span in (True, False):
                for win in (w, None):
                    got = acc.allele_frequency_spectrum(sets, windows=win, mode="branch", polarised=pol, span_normalise=span)
                    want = ts.allele_frequency_spectrum(sets, windows=win, mode="branch", polarised=pol, span_normalise=span)
                    assert got.shape == want.shape
                    assert np.allclose(got, want, rtol=1e-9, atol=1e-12 * np.abs(want).max()), (len(sets), pol, span)
    assert acc.accel_stats["forwarded"] == 0
    # time windows other than [0, inf) (trees.c:3663-3680): every branch is split by time, on the engine
    # that keeps the node of every piece (staged on first use)
    tmax = float(ts.nodes_time.max())
    for tw in ([0, 10.0, np.inf], [0, 3.0, 25.0, 100.0], [0, 0.5 * tmax, tmax, 2 * tmax]):
        for sets in ([s[:50]], [s[:20], s[20:45]]):
            for pol in (False, True):
                for win in (w, None):
                    got = acc.allele_frequency_spectrum(sets, windows=win, mode="branch", time_windows=tw,
                                                        polarised=pol)
                    want = ts.allele_frequency_spectrum(sets, windows=win, mode="branch", time_windows=tw,
                                                        polarised=pol)
                    assert got.shape == want.shape
                    assert np.allclose(got, want, rtol=1e-9, atol=1e-12 * np.abs(want).max()), (tw, len(sets), pol)
    # more than 7 sample sets: the spectrum coordinate travels as one state column
    many = [s[3 * i: 3 * i + 2 + (i % 2)] for i in range(9)]  # 9 sets of 2-3 samples: 3^5 4^4 cells
    for mode in ("site", "branch"):
        for pol in (False, True):
            got = acc.allele_frequency_spectrum(many, windows=w, mode=mode, polarised=pol)
            want = ts.allele_frequency_spectrum(many, windows=w, mode=mode, polarised=pol)
            assert got.shape == want.shape
            assert np.allclose(got, want, rtol=1e-9, atol=1e-12 * np.abs(want).max()), (mode, pol)
    got = acc.allele_frequency_spectrum(many, windows=w, mode="branch", time_windows=[0, 5.0, np.inf])
    want = ts.allele_frequency_spectrum(many, windows=w, mode="branch", time_windows=[0, 5.0, np.inf])
    assert np.allclose(got, want, rtol=1e-9, atol=1e-12 * np.abs(want).max())
    assert acc.accel_stats["forwarded"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["multiroot", "paper", "internal_sample", "unary", "missing"])
def test_branch_afs_fixtures(name):
    from oracle import port
    from tests import fixtures as fx
    from tskit_b200.lowlevel import LLTreeSequence
    t = fx.load(name)
    ll, o = LLTreeSequence(t), port.Oracle(t)
    s = t.samples
    sets = [s[:2], s[2:]]
    sizes = np.array([len(x) for x in sets], dtype=np.uint64)
    flat = np.concatenate(sets).astype(np.int32)
    L = t.sequence_length
    for w in ([0, L], [0, L / 3, L]):
        for pol in (False, True):
            got = ll.allele_frequency_spectrum(sizes, flat, w, [0, np.inf], mode="branch", polarised=pol,
                                               span_normalise=False)
            want = o.branch_allele_frequency_spectrum(sets, windows=w, polarised=pol, span_normalise=False)
            assert np.allclose(got[:, 0], want, rtol=1e-9, atol=1e-12), (name, pol)


def _relvec_cases(ts):
    rng = np.random.default_rng(11)
    n, L = ts.num_samples, ts.sequence_length
    return rng, n, L, ([0, L], np.linspace(0, L, 5), [L / 7, L / 3, 0.9 * L])


def test_oracle_relatedness_vector_pinned_to_reference(ts, wf_small):
    """The oracle's step-by-step restatement of the matvec calculator (trees.c:10445-10816) against the
    reference package: centred and not, windows that do not span the sequence, focal nodes that are
    not samples, several roots (CPU)."""
    from oracle import port
    from tests import fixtures as fx
    o = port.Oracle(wf_small)
    rng, n, L, wins = _relvec_cases(ts)
    for K in (1, 3):
        W = rng.normal(size=(n, K))
        for w in wins:
            for centre in (True, False):
                for span in (True, False):
                    got = o.genetic_relatedness_vector(W, windows=w, centre=centre, span_normalise=span)
                    want = ts.genetic_relatedness_vector(W, windows=w, mode="branch", centre=centre,
                                                         span_normalise=span)
                    assert got.shape == want.shape
                    assert np.allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max()), (K, centre, span)
    nodes = np.array([0, 5, ts.num_nodes - 1, ts.num_nodes // 2, 5], dtype=np.int32)
    W = rng.normal(size=(n, 2))
    got = o.genetic_relatedness_vector(W, windows=wins[2], nodes=nodes, centre=False)
    want = ts.genetic_relatedness_vector(W, windows=wins[2], mode="branch", centre=False, nodes=nodes)
    assert np.allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())
    for name in ("multiroot", "internal_sample", "unary"):
        t = fx.load(name)
        mts = dropin.from_tables(t)
        W = rng.normal(size=(mts.num_samples, 2))
        for centre in (True, False):
            got = port.Oracle(t).genetic_relatedness_vector(W, centre=centre)
            want = mts.genetic_relatedness_vector(W, mode="branch", centre=centre)
            assert np.allclose(got[0], want, rtol=1e-9, atol=1e-12), (name, centre)


@pytest.mark.gpu
def test_relatedness_vector_through_dropin(ts, wf_small):
    """GRM x vector on the device (transposed sweep) == the reference package through the unchanged
    public API, == the oracle; more weight columns than one sweep holds; focal nodes that are not
    samples (node-of-piece engine); the reference's errors."""
    from oracle import port
    acc = dropin.accelerate(ts)
    o = port.Oracle(wf_small)
    rng, n, L, wins = _relvec_cases(ts)
    for K in (1, 3, 11):
        W = rng.normal(size=(n, K))
        for w in wins:
            for centre in (True, False):
                for span in (True, False):
                    got = acc.genetic_relatedness_vector(W, windows=w, mode="branch", centre=centre,
                                                         span_normalise=span)
                    want = ts.genetic_relatedness_vector(W, windows=w, mode="branch", centre=centre,
                                                         span_normalise=span)
                    assert got.shape == want.shape
                    assert np.allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max()), (K, centre, span)
    W = rng.normal(size=(n, 2))
    got = acc.genetic_relatedness_vector(W, windows=wins[1], mode="branch", centre=False)
    want = o.genetic_relatedness_vector(W, windows=wins[1], centre=False)
    assert np.allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())
    # windows=None drops the window axis; a 1-d weight vector is one column
    got = acc.genetic_relatedness_vector(W[:, 0], mode="branch")
    assert got.shape == (n, 1) and np.allclose(got, ts.genetic_relatedness_vector(W[:, 0], mode="branch"), rtol=1e-9,
                                               atol=1e-9 * np.abs(got).max())
    # focal nodes: samples in any order with repeats; then internal nodes
    s = ts.samples()
    for nodes in (s[::-3], np.array([s[4], s[4], s[0]]), np.array([0, 5, ts.num_nodes - 1, ts.num_nodes // 2, 5])):
        got = acc.genetic_relatedness_vector(W, windows=wins[2], mode="branch", centre=False, nodes=nodes)
        want = ts.genetic_relatedness_vector(W, windows=wins[2], mode="branch", centre=False, nodes=nodes)
        assert np.allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())
    assert acc.accel_stats["forwarded"] == 0
    # the matrix-free product agrees with the dense branch GRM of the engine
    G = acc.genetic_relatedness_matrix(mode="branch")
    v = rng.normal(size=n)
    assert np.allclose(acc.genetic_relatedness_vector(v, mode="branch")[:, 0], G @ v, rtol=1e-8,
                       atol=1e-9 * np.abs(G @ v).max())
    # errors, as the reference: site and node mode are refused, bad windows, bad nodes
    for bad in ("site", "node"):
        with pytest.raises(tskit.LibraryError) as e1:
            ts.genetic_relatedness_vector(W, mode=bad)
        with pytest.raises(tskit.LibraryError) as e2:
            acc.genetic_relatedness_vector(W, mode=bad)
        assert str(e1.value) == str(e2.value)
    for kw in (dict(windows=[0, L + 1]), dict(windows=[0, L / 2, L / 2, L]),
               dict(nodes=[ts.num_nodes], centre=False), dict(nodes=[-1], centre=False)):
        with pytest.raises(tskit.LibraryError) as e1:
            ts.genetic_relatedness_vector(W, mode="branch", **kw)
        with pytest.raises(tskit.LibraryError) as e2:
            acc.genetic_relatedness_vector(W, mode="branch", **kw)
        assert str(e1.value) == str(e2.value), kw


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["multiroot", "paper", "internal_sample", "unary", "missing"])
def test_relatedness_vector_fixtures(name):
    from oracle import port
    from tests import fixtures as fx
    from tskit_b200.lowlevel import LLTreeSequence
    t = fx.load(name)
    ll, o = LLTreeSequence(t), port.Oracle(t)
    rng = np.random.default_rng(5)
    W = rng.normal(size=(t.num_samples, 2))
    L = t.sequence_length
    for w in ([0, L], [0, L / 3, L], [L / 4, L / 2]):
        for centre in (True, False):
            got = ll.genetic_relatedness_vector(W, w, mode="branch", centre=centre, nodes=t.samples)
            want = o.genetic_relatedness_vector(W, windows=w, centre=centre)
            assert np.allclose(got, want, rtol=1e-9, atol=1e-12), (name, centre)
        nodes = np.arange(t.num_nodes, dtype=np.int32)[::-1]
        got = ll.genetic_relatedness_vector(W, w, mode="branch", centre=False, nodes=nodes)
        want = o.genetic_relatedness_vector(W, windows=w, centre=False, nodes=nodes)
        assert np.allclose(got, want, rtol=1e-9, atol=1e-12), name


@pytest.mark.gpu
def test_pca_iterates_on_the_engine(ts):
    """`TreeSequence.pca` (trees.py:9284-9557) is a randomised SVD whose operator is the relatedness
    vector product: under the drop-in every product runs on the device, and the result is the
    reference's (same seed; eigenvectors up to sign)."""
    acc = dropin.accelerate(ts)
    L = ts.sequence_length
    for kw in (dict(num_components=3), dict(num_components=2, windows=[0, L / 2, L], samples=ts.samples()[:50]),
               dict(num_components=2, centre=False)):
        want = ts.pca(random_seed=7, **kw)
        before = acc.accel_stats["accelerated"]
        got = acc.pca(random_seed=7, **kw)
        assert acc.accel_stats["accelerated"] > before and acc.accel_stats["forwarded"] == 0
        assert got.factors.shape == want.factors.shape
        assert np.allclose(got.eigenvalues, want.eigenvalues, rtol=1e-7)
        f, g = got.factors.reshape(-1, *got.factors.shape[-2:]), want.factors.reshape(-1, *want.factors.shape[-2:])
        for a, b in zip(f, g):
            sign = np.sign(np.sum(a * b, axis=0))
            assert np.allclose(a * sign, b, atol=1e-6), kw


@pytest.mark.gpu
def test_variants_and_haplotypes_on_the_device_decode(ts):
    """TreeSequence.variants / haplotypes (trees.py:5288-5560) through the device decode: real
    tskit.Variant objects, the reference's genotypes, alleles and missing-data conventions."""
    from tests import fixtures as fx
    acc = dropin.accelerate(ts)
    s = ts.samples()
    L = ts.sequence_length
    cases = [dict(), dict(samples=s[::3]), dict(isolated_as_missing=False), dict(left=0.2 * L, right=0.7 * L),
             dict(samples=s[[5, 2, 9]], copy=False)]
    for kw in cases:
        got = [(v.site.id, v.alleles, v.genotypes.copy(), v.has_missing_data, v.num_alleles) for v in acc.variants(**kw)]
        want = [(v.site.id, v.alleles, v.genotypes.copy(), v.has_missing_data, v.num_alleles) for v in ts.variants(**kw)]
        assert len(got) == len(want) and len(got) > 0
        for a, b in zip(got, want):
            assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2]) and a[3:] == b[3:], (kw, a[0])
    assert acc.accel_stats["forwarded"] == 0 and acc.accel_stats["accelerated"] > 0
    v = next(acc.variants())
    assert isinstance(v, tskit.Variant) and v.genotypes.dtype == np.int32
    assert v.counts() == next(ts.variants()).counts() and v.states().tolist() == next(ts.variants()).states().tolist()
    with pytest.raises(Exception):
        v.decode(0)  # a copy cannot be decoded again, as in the reference
    assert list(acc.haplotypes()) == list(ts.haplotypes())
    assert list(acc.haplotypes(samples=s[:7], left=0.1 * L, right=0.5 * L)) == \
        list(ts.haplotypes(samples=s[:7], left=0.1 * L, right=0.5 * L))
    # a user-supplied allele coding goes to the reference, visibly
    a = [v.genotypes.copy() for v in acc.variants(alleles=("0", "1"))]
    b = [v.genotypes.copy() for v in ts.variants(alleles=("0", "1"))]
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and acc.accel_stats["forwarded"] == 1
    # reference fixtures: stacked / back mutations, several alleles, isolated samples (missing data)
    for name in ("single_tree", "paper", "missing", "multiroot", "internal_sample"):
        t = fx.load(name)
        if t.num_sites == 0:
            continue
        rts = dropin.from_tables(t)
        racc = dropin.accelerate(rts)
        for iso in (True, False):
            got = [(v.site.id, v.alleles, v.genotypes.tolist()) for v in racc.variants(isolated_as_missing=iso)]
            want = [(v.site.id, v.alleles, v.genotypes.tolist()) for v in rts.variants(isolated_as_missing=iso)]
            assert got == want, (name, iso)
