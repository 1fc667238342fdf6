"""GPU suite, larger inputs: a ~2 M-edge Wright-Fisher ARG generated on the fly (seeded), checked
through the C ABI by size-independent properties -- genome shards add up to the whole, fine
windows add up to coarse ones, a million tiny windows, identities between statistics -- and
against the compiled reference where the CPU finishes in seconds."""
import numpy as np
import pytest

from oracle import ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    from tskit_b200.sim import add_mutations, wright_fisher
    t = wright_fisher(20000, 1000, 2e7, ncross=2, seed=5)
    return add_mutations(t, 200000, seed=9).ensure_derived()


@pytest.fixture(scope="module")
def ll(big):
    from tskit_b200.lowlevel import LLTreeSequence
    return LLTreeSequence(big)


def u64(*a):
    return np.array(a, dtype=np.uint64)


def test_matches_compiled_reference(big, ll):
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    r = ref.RefTreeSequence(big)
    s = big.samples
    n = len(s)
    w = np.linspace(0, big.sequence_length, 1001)
    for mode in ("branch", "site"):
        got = ll.diversity(u64(n), s, windows=w, mode=mode)
        want = r.one_way("diversity", [s], windows=w, mode=mode)
        assert np.allclose(got, want, rtol=1e-9, atol=0), mode
    sets = [s[: n // 3], s[n // 3: n // 2], s[n // 2:]]
    flat = np.concatenate(sets).astype(np.int32)
    idx = np.array([[0, 1, 2], [2, 0, 1]], dtype=np.int32)
    got = ll.Y3(u64(*[len(x) for x in sets]), flat, idx, windows=w, mode="branch")
    want = r.k_way("Y3", sets, idx, windows=w, mode="branch")
    assert np.allclose(got, want, rtol=1e-9, atol=0)


def test_shards_windows_and_identities(big, ll):
    from tskit_b200.lowlevel import LLTreeSequence
    s = big.samples
    n = len(s)
    L = big.sequence_length
    w = np.linspace(0, L, 1001)
    whole = ll.diversity(u64(n), s, windows=w, mode="branch", span_normalise=False)
    # genome shards (cut inside windows) add up to the whole, window by window
    cuts = [0.0, 0.31234 * L, 0.7 * L + 0.5, L]
    parts = sum(LLTreeSequence(big, genome_range=(a, b)).diversity(u64(n), s, windows=w, mode="branch",
                                                                  span_normalise=False)
                for a, b in zip(cuts[:-1], cuts[1:]))
    assert np.allclose(parts, whole, rtol=1e-10, atol=0)
    # fine windows add up to coarse ones; one window is the sum of all
    coarse = ll.diversity(u64(n), s, windows=w[::10], mode="branch", span_normalise=False)
    assert np.allclose(whole.reshape(100, 10).sum(axis=1), coarse[:, 0], rtol=1e-10, atol=0)
    one = ll.diversity(u64(n), s, windows=[0, L], mode="branch", span_normalise=False)
    assert np.isclose(whole.sum(), one[0, 0], rtol=1e-10)
    # a million tiny windows (traversal-seeding stress of BASELINE configs[4]): warp-per-window path
    tiny = np.linspace(0, L, 1_000_001)
    fine = ll.diversity(u64(n), s, windows=tiny, mode="branch", span_normalise=False)
    assert fine.shape == (1_000_000, 1)
    assert np.allclose(fine.reshape(1000, 1000).sum(axis=1), whole[:, 0], rtol=1e-10, atol=0)
    # identities: divergence of a set with itself is its diversity; f2(A, B) is symmetric
    a, b = s[: n // 2], s[n // 2:]
    flat = np.concatenate([a, b]).astype(np.int32)
    dv = ll.divergence(u64(len(a), len(b)), flat, np.array([[0, 0], [1, 1]], dtype=np.int32),
                       windows=w, mode="branch")
    di = ll.diversity(u64(len(a), len(b)), flat, windows=w, mode="branch")
    assert np.allclose(dv, di, rtol=1e-12, atol=0)
    f2 = ll.f2(u64(len(a), len(b)), flat, np.array([[0, 1], [1, 0]], dtype=np.int32), windows=w, mode="branch")
    assert np.allclose(f2[:, 0], f2[:, 1], rtol=1e-9, atol=1e-9 * np.abs(f2).max())


def test_custom_summary_many_windows(big, ll):
    """general_stat with a Python summary function (tabulated over every count on the host, evaluated
    on the device) on 10^5 windows, against the built-in statistic with the same formula."""
    s = big.samples
    n = len(s)
    L = big.sequence_length
    W = np.ones((n, 1))
    f = lambda x: x * (n - x) / (n * (n - 1))  # noqa: E731
    w = np.linspace(0, L, 100_001)
    for mode in ("branch", "site"):
        got = ll.general_stat(W, f, 1, windows=w, mode=mode, polarised=True)
        want = ll.diversity(u64(n), s, windows=w, mode=mode, polarised=True)
        assert np.allclose(got, want, rtol=1e-12, atol=0), mode


def test_branch_divergence_matrix_equals_statistic(big, ll):
    s = big.samples
    sets = [s[:5000], s[5000:5001], s[6000:12000], s[12000:12040], s[15000:]]
    sizes = u64(*[len(x) for x in sets])
    flat = np.concatenate(sets).astype(np.int32)
    w = np.array([0.0, 0.25, 0.6]) * big.sequence_length  # not covering the genome
    got = ll.divergence_matrix(w, sample_sets=flat, sample_set_sizes=sizes, mode="branch")
    full = np.array([0.0, 0.25, 0.6, 1.0]) * big.sequence_length
    idx = np.array([[i, j] for i in range(5) for j in range(5)], dtype=np.int32)
    st = ll.divergence(sizes, flat, idx, windows=full, mode="branch").reshape(3, 5, 5)[:2]
    st[:, 1, 1] = 0  # singleton set: the matrix has 0 where the statistic is 0/0 (trees.c:8888-8891)
    assert np.allclose(got, st, rtol=1e-12, atol=0)
