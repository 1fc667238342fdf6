"""GPU suite (run on the B200 box): the CUDA path, called through the C ABI, against the
oracle on identical inputs.  Integer outputs bit-exact; fp64 statistics within rtol 1e-9
(north_star), with an absolute floor of 1e-9 x the largest magnitude in the result for
statistics that are differences of large terms (f2/f3/f4, relatedness)."""
import numpy as np
import pytest

from oracle import port, ref
from tests import fixtures as fx
from tests import plan_model

pytestmark = pytest.mark.gpu

RTOL = 1e-9
ONE_WAY = ["diversity", "segregating_sites", "Y1"]
K_WAY = {"divergence": [[0, 1], [1, 2], [0, 0]], "Y2": [[0, 1], [2, 1]], "f2": [[0, 1], [0, 2]],
         "genetic_relatedness": [[0, 1], [2, 2]], "Y3": [[0, 1, 2]], "f3": [[0, 1, 2], [2, 1, 0]],
         "f4": [[0, 1, 2, 0]]}


def close(got, want, cancelling=False):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape
    assert np.array_equal(np.isnan(got), np.isnan(want))
    atol = RTOL * np.nanmax(np.abs(want), initial=0.0) if cancelling else 0.0
    return np.allclose(got, want, rtol=RTOL, atol=atol, equal_nan=True)


def sets_args(sets):
    sizes = np.array([len(x) for x in sets], dtype=np.uint64)
    flat = np.concatenate([np.asarray(x, dtype=np.int32) for x in sets])
    return sizes, flat


@pytest.fixture(scope="module")
def engines(wf_small):
    from tskit_b200.lowlevel import LLTreeSequence
    return LLTreeSequence(wf_small), port.Oracle(wf_small)


def test_plan_matches_model(wf_small, engines):
    assert plan_model.compare(engines[0], wf_small) == []


def test_plan_matches_model_clipped(wf_small):
    from tskit_b200.lowlevel import LLTreeSequence
    ll = LLTreeSequence(wf_small, genome_range=(30000.0, 70000.0))
    assert plan_model.compare(ll, wf_small, 30000.0, 70000.0) == []


@pytest.mark.parametrize("mode", ["branch", "site"])
@pytest.mark.parametrize("polarised", [False, True])
def test_all_statistics_wright_fisher(wf_small, engines, mode, polarised):
    ll, o = engines
    s = wf_small.samples
    sets = [s[:50], s[50:120], s[120:]]
    sizes, flat = sets_args(sets)
    for windows in (np.array([0.0, wf_small.sequence_length]),
                    np.linspace(0, wf_small.sequence_length, 8),
                    np.array([0.0, 10.5, 11.0, 50000.25, wf_small.sequence_length])):
        for name in ONE_WAY:
            got = getattr(ll, name)(sizes, flat, windows=windows, mode=mode, polarised=polarised)
            want = o.stat(name, sets, windows=windows, mode=mode, polarised=polarised)
            assert close(got, want), (name, mode)
        for name, idx in K_WAY.items():
            got = getattr(ll, name)(sizes, flat, np.array(idx, dtype=np.int32), windows=windows,
                                    mode=mode, polarised=polarised)
            want = o.stat(name, sets, idx, windows=windows, mode=mode, polarised=polarised)
            assert close(got, want, cancelling=True), (name, mode)
        got = ll.genetic_relatedness(sizes, flat, np.array([[0, 1]], dtype=np.int32),
                                     windows=windows, mode=mode, polarised=polarised, centre=False)
        want = o.stat("genetic_relatedness", sets, [[0, 1]], windows=windows, mode=mode,
                      polarised=polarised, centre=False)
        assert close(got, want, cancelling=True)


@pytest.mark.parametrize("variant", ["lane", "c4", "runs", "bins"])
def test_branch_summary_kernel_variants(wf_small, engines, monkeypatch, variant):
    """The default branch summary of a one-column call with uniform windows is the window-run kernel; the
    per-breakpoint delta kernels (one piece per lane; four pieces per thread), the window runs for 2-5
    columns and non-uniform windows (which fall back to the deltas), and the shared-memory window bins
    stay selectable and must agree with the oracle."""
    ll, o = engines
    s = wf_small.samples
    sets = [s[:50], s[50:120], s[120:]]
    sizes, flat = sets_args(sets)
    monkeypatch.setenv("TSKB_SUM_VARIANT", variant)
    for windows in (np.linspace(0, wf_small.sequence_length, 8),
                    np.array([0.0, 10.5, 11.0, 50000.25, wf_small.sequence_length])):
        for polarised in (False, True):
            for name in ONE_WAY:
                got = getattr(ll, name)(sizes, flat, windows=windows, mode="branch", polarised=polarised)
                assert close(got, o.stat(name, sets, windows=windows, mode="branch", polarised=polarised)), name
            for name, idx in K_WAY.items():
                got = getattr(ll, name)(sizes, flat, np.array(idx, dtype=np.int32), windows=windows,
                                        mode="branch", polarised=polarised)
                want = o.stat(name, sets, idx, windows=windows, mode="branch", polarised=polarised)
                assert close(got, want, cancelling=True), name


@pytest.mark.parametrize("mode", ["branch", "site"])
def test_span_normalise_off_and_many_windows(wf_small, engines, mode):
    ll, o = engines
    s = wf_small.samples
    sizes, flat = sets_args([s])
    # windows far smaller than trees: 20 000 (window runs: most pieces cover many windows entirely) and
    # 40 000 (past the limit of the window-run bins: per-breakpoint deltas)
    for count in (20001, 40001):
        windows = np.linspace(0, wf_small.sequence_length, count)
        got = ll.diversity(sizes, flat, windows=windows, mode=mode, span_normalise=False)
        want = o.stat("diversity", [s], windows=windows, mode=mode, span_normalise=False)
        assert close(got, want), count


def test_overlapping_sample_sets_and_eight_sets(wf_small, engines):
    ll, o = engines
    s = wf_small.samples
    sets = [s[i * 20:(i + 2) * 20] for i in range(8)]  # a sample may be in several sets
    sizes, flat = sets_args(sets)
    idx = np.array([(i, j) for i in range(8) for j in range(i, 8)], dtype=np.int32)
    for mode in ("branch", "site"):
        assert close(ll.divergence(sizes, flat, idx, mode=mode, windows=[0, 1e5]),
                     o.stat("divergence", sets, idx, mode=mode))
        assert close(ll.f2(sizes, flat, idx, mode=mode, windows=[0, 1e5]),
                     o.stat("f2", sets, idx, mode=mode), cancelling=True)


@pytest.mark.parametrize("mode", ["branch", "site"])
def test_weighted_statistics(wf_small, engines, mode):
    """fp64-state sweep: trait covariance / correlation and weighted relatedness against the
    oracle's restatement (rtol 1e-9 with an absolute floor: the weights are centred, so states and
    summaries cancel)."""
    from tskit_b200.lowlevel import LibraryError
    ll, o = engines
    n = wf_small.num_samples
    rng = np.random.default_rng(17)
    for K in (1, 2, 3, 7, 13):  # 13 columns: batches of state columns
        W = rng.normal(size=(n, K)) * (1 + np.arange(K)) + np.arange(K)
        for w in (np.array([0.0, wf_small.sequence_length]), np.linspace(0, wf_small.sequence_length, 9)):
            for span in (True, False):
                got = ll.trait_covariance(W, w, mode=mode, span_normalise=span)
                assert close(got, o.trait_covariance(W, windows=w, mode=mode, span_normalise=span), cancelling=True)
                got = ll.trait_correlation(W, w, mode=mode, span_normalise=span)
                assert close(got, o.trait_correlation(W, windows=w, mode=mode, span_normalise=span), cancelling=True)
            if K <= 3:
                Z = np.linalg.qr(np.column_stack([np.ones(n), rng.normal(size=(n, 3))]))[0]
                got = ll.trait_linear_model(W, Z, w, mode=mode, span_normalise=True)
                assert close(got, o.trait_linear_model(W, Z, windows=w, mode=mode), cancelling=True), K
            idx = rng.integers(0, K, size=(9, 2)).astype(np.int32)
            for centre in (True, False):
                for pol in (False, True):
                    got = ll.genetic_relatedness_weighted(W, idx, w, mode=mode, polarised=pol, centre=centre)
                    want = o.genetic_relatedness_weighted(W, idx, windows=w, mode=mode, polarised=pol, centre=centre)
                    assert close(got, want, cancelling=True), (K, centre, pol)
    with pytest.raises(LibraryError) as e:
        ll.trait_covariance(np.zeros((n, 0)), [0, wf_small.sequence_length], mode=mode)
    assert e.value.code == -913
    with pytest.raises(LibraryError) as e:
        ll.trait_covariance(np.zeros((n, 1)), [0, 1.0], mode=mode)
    assert e.value.code == -901
    with pytest.raises(ValueError):
        ll.trait_covariance(np.zeros((n + 1, 1)), [0, wf_small.sequence_length], mode=mode)


def test_many_result_columns(wf_small, engines, monkeypatch):
    """6 or more result columns take the lanes-are-columns branch summary (up to 32 columns per
    pass, so 40 tuples need two passes); both kernels against the oracle."""
    ll, o = engines
    s = wf_small.samples
    sets = [s[25 * i: 25 * i + 20 + i] for i in range(8)]
    sizes, flat = sets_args(sets)
    rng = np.random.default_rng(11)
    for w in (np.linspace(0, wf_small.sequence_length, 12), np.array([0.0, wf_small.sequence_length])):
        pairs = np.array([(i, j) for i in range(8) for j in range(i, 8)], dtype=np.int32)  # 36 columns
        want = o.stat("divergence", sets, pairs, windows=w, mode="branch")
        assert close(ll.divergence(sizes, flat, pairs, windows=w, mode="branch"), want)
        quads = rng.integers(0, 8, size=(40, 4)).astype(np.int32)
        want4 = o.stat("f4", sets, quads, windows=w, mode="branch")
        assert close(ll.f4(sizes, flat, quads, windows=w, mode="branch"), want4, cancelling=True)
        want1 = o.stat("Y1", sets, windows=w, mode="branch", polarised=True)
        assert close(ll.Y1(sizes, flat, windows=w, mode="branch", polarised=True), want1)
        # the calls above fit the shared-memory window bins; the per-breakpoint delta formulation:
        # pieces walked by start breakpoint (default for many columns), in processing order with
        # lanes = columns (the earlier kernel), and one column at a time
        monkeypatch.setenv("TSKB_SUM_VARIANT", "lane")
        for extra in ({}, {"TSKB_COLS_VARIANT": "old"}, {"TSKB_NO_COLS_KERNEL": "1"}):
            for k, v in extra.items():
                monkeypatch.setenv(k, v)
            assert close(ll.divergence(sizes, flat, pairs, windows=w, mode="branch"), want), extra
            assert close(ll.f4(sizes, flat, quads, windows=w, mode="branch"), want4, cancelling=True), extra
            assert close(ll.Y1(sizes, flat, windows=w, mode="branch", polarised=True), want1), extra
            for k in extra:
                monkeypatch.delenv(k)
        monkeypatch.delenv("TSKB_SUM_VARIANT")
    # many columns AND too many windows for the bins: the walk by start breakpoint on its own merits
    w = np.linspace(0, wf_small.sequence_length, 3001)
    pairs = np.array([(i, j) for i in range(8) for j in range(i, 8)], dtype=np.int32)
    for sn in (True, False):
        assert close(ll.divergence(sizes, flat, pairs, windows=w, mode="branch", span_normalise=sn),
                     o.stat("divergence", sets, pairs, windows=w, mode="branch", span_normalise=sn))


@pytest.mark.parametrize("mode", ["branch", "site"])
def test_more_sample_sets_than_one_sweep_carries(wf_small, engines, mode):
    """20 sample sets: result columns are computed in batches of tuples touching <= 8 sets."""
    ll, o = engines
    s = wf_small.samples
    sets = [s[10 * i: 10 * i + 7 + (i % 3)] for i in range(20)]
    sizes, flat = sets_args(sets)
    w = np.linspace(0, wf_small.sequence_length, 6)
    for nm in ONE_WAY:
        assert close(getattr(ll, nm)(sizes, flat, windows=w, mode=mode), o.stat(nm, sets, windows=w, mode=mode))
    rng = np.random.default_rng(3)
    for nm, k in (("divergence", 2), ("Y2", 2), ("f2", 2), ("Y3", 3), ("f3", 3), ("f4", 4)):
        idx = rng.integers(0, 20, size=(37, k)).astype(np.int32)
        got = getattr(ll, nm)(sizes, flat, idx, windows=w, mode=mode)
        assert close(got, o.stat(nm, sets, idx, windows=w, mode=mode), cancelling=True), nm
    idx = np.array([[0, 19], [7, 7]], dtype=np.int32)
    got = ll.genetic_relatedness(sizes, flat, idx, windows=w, mode=mode, centre=False)
    assert close(got, o.stat("genetic_relatedness", sets, idx, windows=w, mode=mode, centre=False), cancelling=True)
    # centred: the mean over ALL 20 sets (trees.c:4729-4753) travels as one extra fp64 state column
    idx = rng.integers(0, 20, size=(23, 2)).astype(np.int32)
    for polarised in (True, False):
        got = ll.genetic_relatedness(sizes, flat, idx, windows=w, mode=mode, centre=True, polarised=polarised)
        want = o.stat("genetic_relatedness", sets, idx, windows=w, mode=mode, centre=True, polarised=polarised)
        assert close(got, want, cancelling=True), polarised


@pytest.mark.parametrize("name", list(fx.ALL))
def test_reference_fixtures(name):
    from tskit_b200.lowlevel import LLTreeSequence
    t = fx.load(name)
    ll, o = LLTreeSequence(t), port.Oracle(t)
    assert plan_model.compare(ll, t) == []
    s = t.samples
    sets = [s[:1], s[1:2], s[2:]]
    sizes, flat = sets_args(sets)
    L = t.sequence_length
    for windows in (np.array([0.0, L]), np.array([0.0, L / 3, L])):
        for mode in ("branch", "site"):
            for pol in (False, True):
                for nm in ONE_WAY:
                    got = getattr(ll, nm)(sizes, flat, windows=windows, mode=mode, polarised=pol)
                    want = o.stat(nm, sets, windows=windows, mode=mode, polarised=pol)
                    assert close(got, want, cancelling=True), (nm, mode, pol)
                for nm, idx in K_WAY.items():
                    got = getattr(ll, nm)(sizes, flat, np.array(idx, dtype=np.int32),
                                          windows=windows, mode=mode, polarised=pol)
                    want = o.stat(nm, sets, idx, windows=windows, mode=mode, polarised=pol)
                    assert close(got, want, cancelling=True), (nm, mode, pol)
    if t.num_edges:
        q = np.array([0.0, L / 2, L * 0.99])
        gp, gc = ll.trees_at(q)
        op, oc = o.trees_at(q)
        assert np.array_equal(gp, op) and np.array_equal(gc, oc)


def test_golden_values():
    from tskit_b200.lowlevel import LLTreeSequence
    ll = LLTreeSequence(fx.load("paper"))
    sizes, flat = sets_args([[0, 1, 2, 3]])
    pi = ll.diversity(sizes, flat, windows=[0, 10], mode="site", span_normalise=False)
    assert abs(pi[0, 0] - fx.PAPER_SITE_DIVERSITY) < 1e-9
    sizes, flat = sets_args([[0]])
    assert np.isnan(ll.diversity(sizes, flat, windows=[0, 10], mode="site")[0, 0])
    assert np.isnan(ll.diversity(sizes, flat, windows=[0, 10], mode="branch")[0, 0])
    ll = LLTreeSequence(fx.load("four_taxa"))
    sizes, flat = sets_args([[0], [1], [2], [3]])
    idx = np.array([[0, 1, 2, 3]], dtype=np.int32)
    got = ll.f4(sizes, flat, idx, windows=fx.FOUR_TAXA_WINDOWS, mode="branch")[:, 0]
    assert np.allclose(got, fx.FOUR_TAXA_F4_0123_WINDOWED, atol=1e-12)
    sizes, flat = sets_args([[0, 1, 2, 3]])
    got = ll.diversity(sizes, flat, windows=fx.FOUR_TAXA_WINDOWS, mode="branch")[:, 0]
    assert np.allclose(got, fx.FOUR_TAXA_DIVERSITY_WINDOWED, atol=1e-12)
    ll = LLTreeSequence(fx.load("case_1"))
    for (i, j), v in fx.CASE_1_BRANCH_DIVERGENCE.items():
        sizes, flat = sets_args([[i], [j]])
        got = ll.divergence(sizes, flat, np.array([[0, 1]], dtype=np.int32), windows=[0, 1.0],
                            mode="branch")
        assert abs(got[0, 0] - v) < 1e-12


def test_error_codes_and_precedence(wf_small, engines):
    from tskit_b200.lowlevel import LibraryError
    ll, _ = engines
    L = wf_small.sequence_length
    s = wf_small.samples
    internal = int(np.nonzero((wf_small.nodes_flags & 1) == 0)[0][0])

    def code(fn):
        with pytest.raises(LibraryError) as e:
            fn()
        return e.value.code

    u64 = lambda *a: np.array(a, dtype=np.uint64)  # noqa: E731
    i32 = lambda *a: np.array(a, dtype=np.int32)  # noqa: E731
    w = [0, L]
    assert code(lambda: ll.diversity(u64(), i32(), windows=w)) == -905
    assert code(lambda: ll.diversity(u64(1, 0), i32(0), windows=w)) == -908
    assert code(lambda: ll.diversity(u64(2), i32(0, 0), windows=w)) == -600
    assert code(lambda: ll.diversity(u64(1), i32(10 ** 6), windows=w)) == -202
    assert code(lambda: ll.diversity(u64(1), i32(internal), windows=w)) == -601
    assert code(lambda: ll.diversity(u64(1), i32(0), windows=[0, L / 2])) == -901
    assert code(lambda: ll.diversity(u64(1), i32(0), windows=[0, L, L / 2])) == -901
    assert code(lambda: ll.diversity(u64(1), i32(0), windows=[1, L])) == -901
    assert code(lambda: ll.divergence(u64(1, 1), i32(0, 1), [[0, 2]], windows=w)) == -907
    # index tuples are checked before sample sets, sample sets before windows
    assert code(lambda: ll.divergence(u64(1, 0), i32(0), [[0, 5]], windows=[0, 1])) == -907
    assert code(lambda: ll.diversity(u64(2), i32(0, 0), windows=[0, 1])) == -600
    assert ll.diversity(u64(1), i32(0), windows=w, mode="node").shape == (len(w) - 1, wf_small.num_nodes, 1)
    with pytest.raises(ValueError):
        ll.diversity(u64(1), i32(0), windows=w, mode="bogus")
    with pytest.raises(ValueError):
        ll.diversity(u64(2), i32(0), windows=w)
    with pytest.raises(ValueError):
        ll.diversity(u64(1), i32(0), windows=[0.0])
    # a same-sample-in-two-sets call is legal (trees.c:2201-2214)
    assert ll.diversity(u64(2, 2), i32(s[0], s[1], s[0], s[2]), windows=w).shape == (1, 2)


def test_time_uncalibrated():
    from tskit_b200.lowlevel import LibraryError, LLTreeSequence
    t = fx.load("paper")
    t.time_uncalibrated = True
    ll = LLTreeSequence(t)
    sizes, flat = sets_args([[0, 1]])
    with pytest.raises(LibraryError) as e:
        ll.diversity(sizes, flat, windows=[0, 10], mode="branch")
    assert e.value.code == -910
    assert ll.diversity(sizes, flat, windows=[0, 10], mode="site").shape == (1, 1)


def test_trees_at_every_breakpoint(wf_small, engines):
    ll, o = engines
    bp = np.unique(np.concatenate([wf_small.edges_left, wf_small.edges_right]))
    q = bp[bp < wf_small.sequence_length][::7]
    gp, gc = ll.trees_at(q)
    op, oc = o.trees_at(q)
    assert np.array_equal(gp, op)
    assert np.array_equal(gc, oc)
    tracked = wf_small.samples[::3]
    gp, gc = ll.trees_at(q[:50], tracked=tracked)
    op, oc = o.trees_at(q[:50], tracked=tracked)
    assert np.array_equal(gp, op) and np.array_equal(gc, oc)


def test_custom_summary_function_tabulated(wf_small, engines):
    """sample_count_stat with a Python f (trees.py:8006-8106) through the device LUT."""
    ll, o = engines
    s = wf_small.samples
    n = len(s)
    W = np.ones((n, 1))

    def f(x):
        return np.array([x[0] * (n - x[0]) / (n * (n - 1)), float(x[0] > 0)])

    windows = np.linspace(0, wf_small.sequence_length, 6)
    for mode in ("branch", "site"):
        for pol in (False, True):
            got = ll.general_stat(W, f, 2, windows=windows, mode=mode, polarised=pol)
            want = o.general_stat(W, f, 2, windows=windows, mode=mode, polarised=pol)
            assert close(got, want), (mode, pol)
    # equals the built-in diversity (up to the unpolarised double count)
    sizes, flat = sets_args([s])
    d = ll.diversity(sizes, flat, windows=windows, mode="branch")
    g = ll.general_stat(W, f, 2, windows=windows, mode="branch")
    assert np.allclose(g[:, 0], d[:, 0], rtol=1e-12)


def test_custom_summary_that_is_nan_only_at_the_roots(wf_small, engines):
    """The reference's running sum takes 0 x f(state) from nodes without a branch above them as well
    (trees.c:1339-1350): a summary that is NaN (or inf) only at the state of a root still turns the
    windows NaN.  The default plan drops those pieces, so such calls run on the plan that keeps every
    piece; results equal the oracle's restatement of the reference loop."""
    ll, o = engines
    s = wf_small.samples
    n = len(s)
    windows = np.linspace(0, wf_small.sequence_length, 5)

    def f_nan(x):   # NaN where a node has every sample below it, as only roots do
        return np.array([np.nan if x[0] == n else x[0] * (n - x[0])])

    def f_inf(x):
        return np.array([np.inf if x[0] == n else float(x[0])])

    W1 = np.ones((n, 1))
    for f in (f_nan, f_inf):
        for pol in (True, False):
            got = ll.general_stat(W1, f, 1, windows=windows, mode="branch", polarised=pol)
            want = o.general_stat(W1, f, 1, windows=windows, mode="branch", polarised=pol)
            assert close(got, want), (f.__name__, pol, got, want)
    # two state columns: through the callback entry point
    W2 = np.ones((n, 2))
    W2[: n // 2, 1] = 0

    def f2(x):
        return np.array([np.nan if x[0] == n else x[0] + x[1]])

    got = ll.general_stat(W2, f2, 1, windows=windows, mode="branch", polarised=True)
    want = o.general_stat(W2, f2, 1, windows=windows, mode="branch", polarised=True)
    assert close(got, want)


def test_general_stat_with_a_callback(wf_small, engines):
    """tsk_treeseq_general_stat with a Python summary over several state columns and over float
    weights (trees.py:7917-8004): the engine sweeps, collects the distinct state vectors on the device and
    calls f once per vector.  Against the oracle's restatement, which calls f at every node update."""
    ll, o = engines
    s = wf_small.samples
    n = len(s)
    rng = np.random.default_rng(5)
    windows = np.linspace(0, wf_small.sequence_length, 6)
    # two overlapping sample sets as 0/1 columns: divergence + a non-linear column
    W2 = np.zeros((n, 2))
    W2[:120, 0] = 1
    W2[80:, 1] = 1
    n0, n1 = W2.sum(axis=0)
    calls = [0]

    def f2(x):
        calls[0] += 1
        return np.array([x[0] * (n1 - x[1]) / (n0 * n1), float(x[0] > 0) * x[1] ** 2, x[0] + 2 * x[1]])

    for mode in ("branch", "site"):
        for pol in (False, True):
            for sn in (True, False):
                got = ll.general_stat(W2, f2, 3, windows=windows, mode=mode, polarised=pol, span_normalise=sn)
                want = o.general_stat(W2, f2, 3, windows=windows, mode=mode, polarised=pol, span_normalise=sn)
                assert close(got, want, cancelling=True), (mode, pol, sn)
    # the callback runs once per distinct vector, not once per node update
    calls[0] = 0
    ll.general_stat(W2, f2, 3, windows=windows, mode="branch", polarised=True)
    distinct_calls = calls[0]
    calls[0] = 0
    o.general_stat(W2, f2, 3, windows=windows, mode="branch", polarised=True)
    assert 0 < distinct_calls < calls[0] / 5
    # float weights, three columns; many result columns (the by-column summary path)
    Wf = rng.normal(size=(n, 3))

    def f3(x):
        return np.concatenate([x * x, [x[0] * x[1], np.abs(x).sum(), 1.0, x[2] - x[0]]])

    for mode in ("branch", "site"):
        got = ll.general_stat(Wf, f3, 7, windows=windows, mode=mode, polarised=False)
        want = o.general_stat(Wf, f3, 7, windows=windows, mode=mode, polarised=False)
        assert np.allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max()), mode
    # failures of the callback surface as they are
    with pytest.raises(ZeroDivisionError):
        ll.general_stat(W2, lambda x: np.array([1 / 0]), 1, windows=windows, mode="branch")
    with pytest.raises(ValueError, match="wrong dimension"):
        ll.general_stat(W2, lambda x: np.array([1.0, 2.0]), 1, windows=windows, mode="branch")


def test_genome_range_shards_sum_to_whole(wf_small, engines):
    """multi-GPU decomposition: shards over [a, b) computed independently add up per window"""
    from tskit_b200.lowlevel import LLTreeSequence
    ll, o = engines
    s = wf_small.samples
    sizes, flat = sets_args([s[:100], s[100:]])
    idx = np.array([[0, 1]], dtype=np.int32)
    L = wf_small.sequence_length
    w = np.linspace(0, L, 10)
    cuts = [0.0, 21000.5, w[4], 77777.0, L]  # inside windows and on a window edge
    for mode in ("branch", "site"):
        whole = ll.divergence(sizes, flat, idx, windows=w, mode=mode, span_normalise=False)
        acc = np.zeros_like(whole)
        for a, b in zip(cuts[:-1], cuts[1:]):
            part = LLTreeSequence(wf_small, genome_range=(a, b))
            acc += part.divergence(sizes, flat, idx, windows=w, mode=mode, span_normalise=False)
        assert np.allclose(acc, whole, rtol=1e-11)
        assert close(whole, o.stat("divergence", [s[:100], s[100:]], idx, windows=w, mode=mode,
                                   span_normalise=False))
    # many columns and more windows (the window-run kernels keep bins only for the windows a range meets)
    sets8 = [s[25 * i: 25 * i + 22] for i in range(8)]
    sizes8, flat8 = sets_args(sets8)
    pairs = np.array([(i, j) for i in range(8) for j in range(i + 1, 8)], dtype=np.int32)
    for w in (np.linspace(0, L, 41), np.array([0.0, 500.0, 21000.5, 30000.0, 77777.0, 90000.0, L])):
        whole = ll.divergence(sizes8, flat8, pairs, windows=w, mode="branch", span_normalise=False)
        one = ll.diversity(sizes8[:1], flat8[:22], windows=w, mode="branch", span_normalise=False)
        acc, acc1 = np.zeros_like(whole), np.zeros_like(one)
        for a, b in zip(cuts[:-1], cuts[1:]):
            part = LLTreeSequence(wf_small, genome_range=(a, b))
            acc += part.divergence(sizes8, flat8, pairs, windows=w, mode="branch", span_normalise=False)
            acc1 += part.diversity(sizes8[:1], flat8[:22], windows=w, mode="branch", span_normalise=False)
        assert np.allclose(acc, whole, rtol=1e-11)
        assert np.allclose(acc1, one, rtol=1e-11)
        assert close(whole, o.stat("divergence", sets8, pairs, windows=w, mode="branch", span_normalise=False))
        assert close(one, o.stat("diversity", sets8[:1], windows=w, mode="branch", span_normalise=False))


def test_relatedness_vector_genome_range_shards_sum_to_whole(wf_small, engines):
    """the relatedness vector over [a, b) shards: un-normalised rows add up (also when centred: centring
    the output rows is linear), which is what tskit_b200.sharding sums over the ranks"""
    from tskit_b200.lowlevel import LLTreeSequence
    ll, o = engines
    s = wf_small.samples
    L = wf_small.sequence_length
    w = np.linspace(0, L, 6)
    cuts = [0.0, 21000.5, w[2], 77777.0, L]
    wt = np.random.default_rng(2).normal(size=(len(s), 3))
    for centre in (True, False):
        whole = ll.genetic_relatedness_vector(wt, w, mode="branch", span_normalise=False, centre=centre, nodes=s)
        acc = np.zeros_like(whole)
        for a, b in zip(cuts[:-1], cuts[1:]):
            part = LLTreeSequence(wf_small, genome_range=(a, b))
            acc += part.genetic_relatedness_vector(wt, w, mode="branch", span_normalise=False, centre=centre,
                                                   nodes=s)
        assert np.allclose(acc, whole, rtol=1e-9, atol=1e-9 * np.abs(whole).max())
        want = o.genetic_relatedness_vector(wt, windows=w, centre=centre, span_normalise=False)
        assert np.allclose(whole, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())


def test_relatedness_vector_golden(wf_small, engines):
    """the device path against the committed vectors of the reference package (tests/golden/)"""
    from tests.test_oracle import relvec_golden_cases, relvec_golden_weights
    ll, _ = engines
    W = relvec_golden_weights(wf_small.num_samples)
    for key, kw, want in relvec_golden_cases():
        nodes = wf_small.samples if kw["nodes"] is None else kw["nodes"]
        got = ll.genetic_relatedness_vector(W, kw["windows"], mode="branch", span_normalise=kw["span_normalise"],
                                            centre=kw["centre"], nodes=nodes)
        assert got.shape == want.shape
        assert np.allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max()), key


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_against_compiled_reference_1k(wf_1k):
    from tskit_b200.lowlevel import LLTreeSequence
    ll, r = LLTreeSequence(wf_1k), ref.RefTreeSequence(wf_1k)
    s = wf_1k.samples
    sets = [s[:300], s[300:]]
    sizes, flat = sets_args(sets)
    w = np.linspace(0, wf_1k.sequence_length, 101)
    for mode in ("branch", "site"):
        assert close(ll.diversity(sizes, flat, windows=w, mode=mode),
                     r.one_way("diversity", sets, windows=w, mode=mode))
        assert close(ll.divergence(sizes, flat, np.array([[0, 1]], dtype=np.int32), windows=w,
                                   mode=mode),
                     r.k_way("divergence", sets, [[0, 1]], windows=w, mode=mode))


def test_concurrent_calls_and_engine_lifecycle(wf_small):
    """The C library is re-entrant on a const tree sequence; the engine serialises calls per handle
    (mutex) and runs handles side by side.  Threads hammer two engines with different statistics;
    every result must equal the serial one.  Then engines are created and freed repeatedly and
    the device memory must come back."""
    import threading
    import torch
    from tskit_b200.lowlevel import LLTreeSequence
    a, b = LLTreeSequence(wf_small), LLTreeSequence(wf_small)
    s = wf_small.samples
    sets = [s[:70], s[70:]]
    sizes, flat = sets_args(sets)
    w = np.linspace(0, wf_small.sequence_length, 33)
    idx = np.array([[0, 1]], dtype=np.int32)
    jobs = [
        lambda e: e.diversity(sizes, flat, windows=w, mode="branch"),
        lambda e: e.divergence(sizes, flat, idx, windows=w, mode="site"),
        lambda e: e.f2(sizes, flat, idx, windows=w, mode="branch"),
        lambda e: e.divergence_matrix(w[:5], mode="site"),
        lambda e: e.diversity(sizes, flat, windows=w, mode="node"),
    ]
    want = [job(a) for job in jobs]
    errors = []

    def worker(k):
        try:
            for r in range(6):
                j = (k + r) % len(jobs)
                got = jobs[j](a if (k + r) % 2 else b)
                if not np.allclose(got, want[j], rtol=1e-9, atol=1e-9 * np.abs(want[j]).max(), equal_nan=True):
                    errors.append((k, r, j))
        except Exception as e:  # noqa: BLE001
            errors.append((k, repr(e)))
    threads = [threading.Thread(target=worker, args=(k,)) for k in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert errors == []
    a.close()
    b.close()
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(4):
        e = LLTreeSequence(wf_small)
        e.diversity(sizes, flat, windows=w, mode="branch")
        e.diversity(sizes, flat, windows=w, mode="node")
        e.close()
    torch.cuda.synchronize()
    assert abs(torch.cuda.mem_get_info()[0] - free0) < 64 << 20


def test_stacked_multiallelic_mutations(wf_small):
    """Several mutations per site on nested and unrelated nodes, four derived states, back
    mutations to the ancestral state: the general per-site allele path, the decode's overwrite
    order and the one-hot tensor-core contraction with more than two alleles, against the oracle."""
    import copy
    from tskit_b200.lowlevel import LLTreeSequence
    t = copy.deepcopy(wf_small)
    rng = np.random.default_rng(23)
    L = t.sequence_length
    pos = np.sort(rng.choice(int(L) - 1, size=400, replace=False)).astype(np.float64) + 0.5
    msite, mnode, mstate = [], [], []
    for j, x in enumerate(pos):
        alive = np.nonzero((t.edges_left <= x) & (t.edges_right > x))[0]
        k = int(rng.integers(1, 6))
        nodes = rng.choice(np.unique(np.concatenate([t.edges_child[alive], t.edges_parent[alive]])), size=k,
                           replace=False)  # one mutation per node and site: table order is unambiguous
        nodes = nodes[np.argsort(-t.nodes_time[nodes], kind="stable")]  # older first: parents precede
        for u in nodes:
            msite.append(j); mnode.append(int(u)); mstate.append(int(rng.integers(0, 4)))
    S, Mu = len(pos), len(msite)
    t.sites_position = pos
    t.sites_ancestral_state = np.full(S, ord("0"), dtype=np.int8)
    t.sites_ancestral_state_offset = np.arange(S + 1, dtype=np.uint64)
    t.mutations_site = np.array(msite, dtype=np.int32)
    t.mutations_node = np.array(mnode, dtype=np.int32)
    t.mutations_derived_state = (np.array(mstate) + ord("0")).astype(np.int8)
    t.mutations_derived_state_offset = np.arange(Mu + 1, dtype=np.uint64)
    t.mutations_parent = None
    t.ensure_derived()
    ll, o = LLTreeSequence(t), port.Oracle(t)
    for missing in (True, False):
        assert np.array_equal(ll.genotype_matrix(isolated_as_missing=missing),
                              o.genotype_matrix(isolated_as_missing=missing))
    if ref.available():  # the oracle itself against the compiled reference on this input
        assert np.array_equal(o.genotype_matrix(), ref.RefTreeSequence(t).genotype_matrix())
    s = t.samples
    sets = [s[:60], s[60:61], s[61:]]
    sizes, flat = sets_args(sets)
    w = np.linspace(0, L, 7)
    for pol in (False, True):
        for nm in ONE_WAY:
            assert close(getattr(ll, nm)(sizes, flat, windows=w, mode="site", polarised=pol),
                         o.stat(nm, sets, windows=w, mode="site", polarised=pol), cancelling=True), nm
        for nm, idx in K_WAY.items():
            got = getattr(ll, nm)(sizes, flat, np.array(idx, dtype=np.int32), windows=w, mode="site", polarised=pol)
            assert close(got, o.stat(nm, sets, idx, windows=w, mode="site", polarised=pol), cancelling=True), nm
    got = ll.divergence_matrix(w, mode="site", span_normalise=False)
    assert np.array_equal(got, o.divergence_matrix(None, windows=w, mode="site", span_normalise=False))


def test_init_rejects_malformed_tables(wf_small):
    """tskb_treeseq_init replaces tsk_treeseq_init: tables that tsk_table_collection_check_integrity
    (tables.c:10362-10640, 10894-10930) refuses are refused with the same code, before any row is
    used as an index on the host or the device."""
    import copy
    from tskit_b200.lowlevel import LibraryError, LLTreeSequence

    def broken(**cols):
        t = copy.copy(wf_small)
        for k, v in cols.items():
            setattr(t, k, v)
        return t

    def poke(col, i, v):
        a = getattr(wf_small, col).copy()
        a[i] = v
        return {col: a}

    E, N, S, Mu = wf_small.num_edges, wf_small.num_nodes, wf_small.num_sites, wf_small.num_mutations
    assert S > 3 and Mu > 3
    cases = [
        (poke("edges_parent", 5, N), -202), (poke("edges_parent", 5, -1), -300),
        (poke("edges_child", 7, -1), -301), (poke("edges_child", 7, N + 3), -202),
        (poke("edges_left", 3, -1.0), -310), (poke("edges_right", 3, wf_small.sequence_length + 1), -309),
        (poke("edges_left", 9, np.inf), -211),
        (poke("edges_right", 11, wf_small.edges_left[11]), -307),
        (poke("edges_child", 2, int(wf_small.edges_parent[2])), -306),
        (poke("nodes_time", N - 1, np.nan), -210),
        (poke("edge_insertion_order", 4, E), -203), (poke("edge_removal_order", 4, -2), -203),
        (poke("sites_position", 2, -3.0), -402),
        (poke("sites_position", 2, float(wf_small.sites_position[1])), -401),
        (poke("sites_position", 2, float(wf_small.sites_position[0])), -400),
        (poke("mutations_site", 1, S), -205), (poke("mutations_node", 1, N), -202),
        (poke("mutations_parent", 1, Mu), -206), (poke("mutations_parent", 1, 1), -501),
        (poke("mutations_parent", 1, 3), -502),
        (poke("mutations_derived_state_offset", 2, 10**9), -200),
    ]
    for cols, code in cases:
        with pytest.raises(LibraryError) as e:
            LLTreeSequence(broken(**cols))
        assert e.value.code == code, (list(cols), e.value.code, code)
    # the first failing row of the first failing table wins, as in the reference's loop
    both = {**poke("edges_parent", 8, N), **poke("mutations_node", 0, -5)}
    with pytest.raises(LibraryError) as e:
        LLTreeSequence(broken(**both))
    assert e.value.code == -202
    LLTreeSequence(wf_small).close()  # and the device is still usable afterwards


def test_restricted_range_engines_on_a_repeated_genome(wf_small):
    """What bench.py --gpus N does on every rank, here one range after the other on one GPU: the
    genome repeated 3 x (sim.repeat_genome), cut by sharding.plan_shards, every range staged from the
    restricted tables (sharding.restrict_tables) and evaluated with inputs and partials in HBM; the sum
    of the partials equals the oracle on the whole tables, and the host path of the sharded wrapper
    (pinned staging, device-side normalisation) equals it too."""
    import torch
    from tskit_b200 import sharding
    from tskit_b200.lowlevel import STAT_BRANCH, STAT_SITE, STAT_SPAN_NORMALISE, LLTreeSequence
    from tskit_b200.sim import repeat_genome
    t = repeat_genome(wf_small, 3)
    W = 12
    windows = np.linspace(0, t.sequence_length, W + 1)
    o = port.Oracle(t)
    s = t.samples
    sets = [s[:70], s[70:]]
    sizes, flat = sets_args(sets)
    idx = np.array([[0, 1], [1, 1]], dtype=np.int32)
    d_sets = torch.from_numpy(flat).cuda()
    ranges = sharding.plan_shards(t, windows, 4)
    for mode, flag in (("branch", STAT_BRANCH), ("site", STAT_SITE)):
        total = torch.zeros((W, 2), dtype=torch.float64, device="cuda")
        part = torch.empty_like(total)
        for rng in ranges:
            local = sharding.restrict_tables(t, *rng)
            ll = LLTreeSequence(local, genome_range=rng)
            ll.stat_device("divergence", sizes, d_sets.data_ptr(), idx, windows, flag, part.data_ptr())
            total += part
            ll.close()
        want = o.stat("divergence", sets, idx, windows=windows, mode=mode, span_normalise=False)
        assert close(total.cpu().numpy(), want)
        sh = sharding.ShardedTreeSequence(t, windows, 0, 1)
        got = sh.stat_host("divergence", sizes, flat, idx, windows, flag | STAT_SPAN_NORMALISE)
        assert close(got, o.stat("divergence", sets, idx, windows=windows, mode=mode))


def test_peer_exchange_single_rank(wf_small, engines):
    """tskb_exchange_sum with a world of one: the push / signal / wait / sum kernels on the local
    buffers, both parities, with and without span normalisation along either axis (the multi-rank
    path is exercised by bench.py --gpus N, whose result is checked against the reference)."""
    import torch
    from tskit_b200 import sharding
    ll, _ = engines
    ex = sharding.PeerExchange(ll, 64, 0, 1)
    w = np.array([0.0, 1.0, 3.0, 7.0, 8.0, 20.0, 21.0])
    spans = torch.from_numpy(np.diff(w)).cuda()
    a = torch.arange(1, 13, dtype=torch.float64, device="cuda").reshape(6, 2)
    for _ in range(3):
        out = torch.empty_like(a)
        ex.sum_into(a, out)
        assert torch.equal(out, a)
        ex.sum_into(a, out, w, window_axis=0)
        assert torch.allclose(out, a / spans[:, None], rtol=1e-15, atol=0)
    b = a.t().contiguous()  # statistics stacked along axis 0, windows along axis 1
    ex.sum_into(b, b, w, window_axis=1)
    assert torch.allclose(b, a.t() / spans[None, :], rtol=1e-15, atol=0)
    with pytest.raises(ValueError):
        ex.sum_into(torch.zeros(65, dtype=torch.float64, device="cuda"), torch.zeros(65, dtype=torch.float64, device="cuda"))
    # results too large for the single-CTA kernel: grid-wide push, signal / wait, sum
    big = sharding.PeerExchange(ll, 70000, 0, 1)
    wb = np.concatenate([[0.0], np.cumsum(np.random.default_rng(2).uniform(0.5, 2.0, 17500))])
    xb = torch.from_numpy(np.random.default_rng(3).normal(size=(17500, 4))).cuda()
    for _ in range(3):
        ob = torch.empty_like(xb)
        big.sum_into(xb, ob, wb)
        assert torch.equal(ob, xb / torch.from_numpy(np.diff(wb)).cuda()[:, None])


@pytest.mark.parametrize("world,count", [(2, 96), (4, 5000), (3, 30000)])
def test_peer_exchange_several_ranks_in_one_process(world, count):
    """The exchange protocol with several ranks: `world` exchange objects on this device, connected in
    process (tskb_exchange_connect_local), one host thread per rank, 7 calls each (both parities; the
    single-CTA kernel -- ranks sharing ONE device cannot rely on the grid-wide kernels of different
    ranks being co-scheduled; those run in test_peer_exchange_single_rank and across GPUs in
    bench.py --config c3).  Every rank must hold the rank-ordered sum, bit for
    bit the same, span-normalised."""
    import threading

    import torch
    from tskit_b200 import sharding
    members = [sharding.PeerExchange(None, count, r, world, connect=False) for r in range(world)]
    sharding.PeerExchange.connect_local(members)
    rng = np.random.default_rng(5)
    W = count // 4
    w = np.concatenate([[0.0], np.cumsum(rng.uniform(0.5, 2.0, W))])
    parts = [[torch.from_numpy(rng.normal(size=(W, 4))).cuda() for _ in range(world)] for _ in range(7)]
    # everything is allocated before the ranks start: a cudaMalloc (torch's allocator growing) waits for
    # the device to drain, i.e. for another rank's kernel that is itself waiting for this rank's push --
    # with one process per GPU the ranks do not share a device
    outs = [[torch.empty_like(parts[c][r]) for r in range(world)] for c in range(7)]
    sharding._spans(w, parts[0][0], 0)
    errors = []

    def run(r):
        try:
            for c in range(7):
                members[r].sum_into(parts[c][r], outs[c][r], w if c % 2 else None, wait=(c % 3 != 0))
            members[r].status()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    torch.cuda.synchronize()
    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    spans = torch.from_numpy(np.diff(w)).cuda()[:, None]
    for c in range(7):
        want = parts[c][0].clone()
        for r in range(1, world):
            want += parts[c][r]
        if c % 2:
            want = want / spans
        for r in range(world):
            assert torch.equal(outs[c][r], want), (c, r)
    for m in members:
        m.close()
