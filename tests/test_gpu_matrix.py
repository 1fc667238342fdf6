"""GPU suite, matrix path: genotype decode (bit-exact against the oracle and the reference's
golden vectors) and the site-mode divergence matrix (exact integer counts, so fp64 results are
compared at 1e-12), through the C ABI."""
import numpy as np
import pytest

from oracle import port
from tests import fixtures as fx

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engines(wf_small):
    from tskit_b200.lowlevel import LLTreeSequence
    return LLTreeSequence(wf_small), port.Oracle(wf_small)


def test_genotypes_golden_single_tree():
    from tskit_b200.lowlevel import LLTreeSequence
    ll = LLTreeSequence(fx.load("single_tree"))
    assert np.array_equal(ll.genotype_matrix(), fx.SINGLE_TREE_GENOTYPES)  # test_genotypes.c:447-503


@pytest.mark.parametrize("name", [n for n in fx.ALL if n != "empty"])
def test_genotypes_reference_fixtures(name):
    from tskit_b200.lowlevel import LLTreeSequence
    t = fx.load(name)
    if t.num_sites == 0:
        pytest.skip("no sites")
    ll, o = LLTreeSequence(t), port.Oracle(t)
    for missing in (True, False):
        assert np.array_equal(ll.genotype_matrix(isolated_as_missing=missing),
                              o.genotype_matrix(isolated_as_missing=missing)), (name, missing)
    sub = t.samples[::2]
    assert np.array_equal(ll.genotype_matrix(samples=sub), o.genotype_matrix(samples=sub))


def test_genotypes_wright_fisher(wf_small, engines):
    ll, o = engines
    g = ll.genotype_matrix()
    assert g.dtype == np.int8 and g.shape == (wf_small.num_sites, wf_small.num_samples)
    assert np.array_equal(g, o.genotype_matrix())
    sub = wf_small.samples[[5, 3, 100, 7]]  # any order, any subset
    assert np.array_equal(ll.genotype_matrix(samples=sub, isolated_as_missing=False),
                          o.genotype_matrix(samples=sub, isolated_as_missing=False))


def test_genotypes_1k(wf_1k):
    from tskit_b200.lowlevel import LLTreeSequence
    ll, o = LLTreeSequence(wf_1k), port.Oracle(wf_1k)
    assert np.array_equal(ll.genotype_matrix(), o.genotype_matrix())


def test_divmat_golden_single_tree():
    from tskit_b200.lowlevel import LLTreeSequence
    ll = LLTreeSequence(fx.load("single_tree"))
    d = ll.divergence_matrix([0, 1.0], mode="site", span_normalise=False)
    assert np.array_equal(d[0], fx.SINGLE_TREE_D_SITE)  # test_stats.c:1459-1473


def test_divmat_site_wright_fisher(wf_small, engines):
    ll, o = engines
    L = wf_small.sequence_length
    s = wf_small.samples
    for windows in ([0, L], np.linspace(0, L, 5), [L * 0.1, L * 0.35, L * 0.8]):
        for span in (True, False):
            got = ll.divergence_matrix(windows, mode="site", span_normalise=span)
            want = o.divergence_matrix(None, windows=windows, mode="site", span_normalise=span)
            assert got.shape == want.shape
            assert np.allclose(got, want, rtol=1e-12, atol=0)
    sets = [s[:40], s[40:41], s[50:170]]
    sizes = np.array([len(x) for x in sets], dtype=np.uint64)
    flat = np.concatenate(sets).astype(np.int32)
    w = np.linspace(0, L, 4)
    got = ll.divergence_matrix(w, sample_sets=flat, sample_set_sizes=sizes, mode="site")
    want = o.divergence_matrix(sets, windows=w, mode="site")
    assert np.allclose(got, want, rtol=1e-12, atol=0)
    assert np.all(got[:, 1, 1] == 0)  # singleton set: diagonal 0, not NaN (trees.c:8888-8891)


def test_divmat_site_1k_tensor_core_paths(wf_1k, monkeypatch):
    """1000 samples = 8 x 8 blocks of 128 with a ragged last block, k-ranges that start and end
    inside 16-byte chunks; the biallelic tcgen05 Gram kernel (256 x 256 tiles), the tcgen05 one-hot
    kernel and the legacy mma.sync kernel must all give the exact integer counts of the oracle."""
    from tskit_b200.lowlevel import LLTreeSequence
    ll, o = LLTreeSequence(wf_1k), port.Oracle(wf_1k)
    L = wf_1k.sequence_length
    windows = [0, L * 0.013, L * 0.4, L * 0.41, L]
    want = o.divergence_matrix(None, windows=windows, mode="site", span_normalise=False)
    got = ll.divergence_matrix(windows, mode="site", span_normalise=False)
    assert np.array_equal(got, want)
    for mode in ("onehot", "legacy"):  # tcgen05 one-hot kernel (any alleles), mma.sync kernel
        monkeypatch.setenv("TSKB_MATRIX", mode)
        got = ll.divergence_matrix(windows, mode="site", span_normalise=False)
        assert np.array_equal(got, want), mode


@pytest.mark.parametrize("name", ["paper", "nonbinary", "multiroot", "missing", "case_1"])
def test_divmat_site_fixtures(name):
    from tskit_b200.lowlevel import LLTreeSequence
    t = fx.load(name)
    ll, o = LLTreeSequence(t), port.Oracle(t)
    L = t.sequence_length
    got = ll.divergence_matrix([0, L / 2, L], mode="site")
    want = o.divergence_matrix(None, windows=[0, L / 2, L], mode="site")
    assert np.allclose(got, want, rtol=1e-12, atol=0)


def test_divmat_errors(wf_small, engines):
    from tskit_b200.lowlevel import LibraryError
    ll, _ = engines
    L = wf_small.sequence_length
    s = wf_small.samples

    def code(fn):
        with pytest.raises(LibraryError) as e:
            fn()
        return e.value.code

    u64 = lambda *a: np.array(a, dtype=np.uint64)  # noqa: E731
    i32 = lambda *a: np.array(a, dtype=np.int32)  # noqa: E731
    assert code(lambda: ll.divergence_matrix([0, L], mode="node")) == -909
    assert code(lambda: ll.divergence_matrix([0, L + 1])) == -901
    assert code(lambda: ll.divergence_matrix([-1, L])) == -901
    assert code(lambda: ll.divergence_matrix([0, L, L / 2])) == -901
    assert code(lambda: ll.divergence_matrix([0, L], sample_sets=i32(s[0], s[0]), sample_set_sizes=u64(1, 1))) == -600
    internal = int(np.nonzero((wf_small.nodes_flags & 1) == 0)[0][0])
    assert code(lambda: ll.divergence_matrix([0, L], sample_sets=i32(internal), sample_set_sizes=u64(1))) == -601
    assert code(lambda: ll.divergence_matrix([0, L], sample_sets=i32(10 ** 7), sample_set_sizes=u64(1))) == -202


def test_divmat_branch_golden_single_tree():
    from tskit_b200.lowlevel import LLTreeSequence
    ll = LLTreeSequence(fx.load("single_tree"))
    d = ll.divergence_matrix([0, 1.0], mode="branch", span_normalise=False)
    assert np.allclose(d[0], fx.SINGLE_TREE_D_BRANCH, rtol=1e-12, atol=0)  # test_stats.c:1459-1473


def test_divmat_branch_wright_fisher(wf_small, engines):
    """Branch-mode matrix (trees.c:8579-8676) = the sweep engine's branch divergence over blocks of
    sample sets; compared with the oracle's per-tree MRCA definition."""
    ll, o = engines
    L = wf_small.sequence_length
    s = wf_small.samples
    sub = s[::13].astype(np.int32)  # 24 singleton sets: several 4 x 4 blocks, a ragged last one
    ones = np.ones(len(sub), dtype=np.uint64)
    for windows in ([0, L], np.linspace(0, L, 5), [L * 0.1, L * 0.35, L * 0.8]):
        for span in (True, False):
            got = ll.divergence_matrix(windows, sample_sets=sub, sample_set_sizes=ones, mode="branch",
                                       span_normalise=span)
            want = o.divergence_matrix([[u] for u in sub], windows=windows, mode="branch",
                                       span_normalise=span)
            assert got.shape == want.shape
            assert np.allclose(got, want, rtol=1e-9, atol=0)
            assert np.all(got[:, np.arange(len(sub)), np.arange(len(sub))] == 0)
    sets = [s[:40], s[40:41], s[50:120], s[120:130], s[140:200]]
    sizes = np.array([len(x) for x in sets], dtype=np.uint64)
    flat = np.concatenate(sets).astype(np.int32)
    w = np.linspace(0, L, 4)
    got = ll.divergence_matrix(w, sample_sets=flat, sample_set_sizes=sizes, mode="branch")
    want = o.divergence_matrix(sets, windows=w, mode="branch")
    assert np.allclose(got, want, rtol=1e-9, atol=0), (got[0], want[0])
    assert np.all(got[:, 1, 1] == 0)  # singleton set: diagonal 0, not NaN (trees.c:8888-8891)


@pytest.mark.parametrize("name", ["paper", "nonbinary", "multiroot", "internal_sample", "case_1"])
def test_divmat_branch_fixtures(name):
    from tskit_b200.lowlevel import LLTreeSequence
    t = fx.load(name)
    ll, o = LLTreeSequence(t), port.Oracle(t)
    L = t.sequence_length
    got = ll.divergence_matrix([0, L / 2, L], mode="branch")
    want = o.divergence_matrix(None, windows=[0, L / 2, L], mode="branch")
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)


def test_genetic_relatedness_matrix_through_dropin(wf_small):
    tskit = pytest.importorskip("tskit")
    from tskit_b200 import dropin
    ts = dropin.from_tables(wf_small)
    acc = dropin.accelerate(ts)
    s = ts.samples()
    sets = [list(s[:30]), list(s[30:90]), list(s[90:])]
    w = np.linspace(0, ts.sequence_length, 4)
    got = acc.genetic_relatedness_matrix(sample_sets=sets, windows=w, mode="site")
    want = ts.genetic_relatedness_matrix(sample_sets=sets, windows=w, mode="site")
    assert np.allclose(got, want, rtol=1e-10, atol=1e-12 * np.abs(want).max())
    got = acc.divergence_matrix(windows=w, mode="site")
    want = ts.divergence_matrix(windows=w, mode="site")
    assert np.allclose(got, want, rtol=1e-12)
    # branch mode: the reference's per-tree MRCA loop against the sweep engine
    got = acc.genetic_relatedness_matrix(sample_sets=sets, windows=w, mode="branch")
    want = ts.genetic_relatedness_matrix(sample_sets=sets, windows=w, mode="branch")
    assert np.allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())
    some = [[int(u)] for u in s[:11]]
    got = acc.divergence_matrix(sample_sets=some, windows=w, mode="branch")
    want = ts.divergence_matrix(sample_sets=some, windows=w, mode="branch")
    assert np.allclose(got, want, rtol=1e-9)
    assert acc.accel_stats["forwarded"] == 0


def test_matrices_shard_by_genome_range(wf_small):
    """SURVEY 8e rows 2-4 on one GPU, one range after the other: the per-range partial divergence
    matrices (site mode: integer counts, so the sum is exact; branch mode: within rounding) add up to
    the oracle's whole-genome matrix, and the per-range genotype blocks concatenate to the whole decode.
    Ranges are staged from the restricted tables, as the ranks of sharding.ShardedTreeSequence do."""
    from tskit_b200 import sharding
    from tskit_b200.lowlevel import LLTreeSequence
    t = wf_small
    o = port.Oracle(t)
    L = t.sequence_length
    windows = np.array([0.0, 0.2 * L, 0.55 * L, L])
    s = t.samples
    sets = [s[:30], s[30:31], s[40:90], s[100:]]
    sizes = np.array([len(x) for x in sets], dtype=np.uint64)
    flat = np.concatenate(sets).astype(np.int32)
    ranges = sharding.plan_shards(t, np.linspace(0, L, 41), 3)
    for mode in ("site", "branch"):
        total = np.zeros((3, 4, 4))
        total_all = np.zeros((3, 40, 40))
        for rng in ranges:
            ll = LLTreeSequence(sharding.restrict_tables(t, *rng), genome_range=rng)
            total += ll.divergence_matrix(windows, sample_sets=flat, sample_set_sizes=sizes, mode=mode,
                                          span_normalise=False)
            total_all += ll.divergence_matrix(windows, sample_sets=s[:40], sample_set_sizes=np.ones(40, dtype=np.uint64),
                                              mode=mode, span_normalise=False)
            ll.close()
        want = o.divergence_matrix(sets, windows=windows, mode=mode, span_normalise=False)
        want_all = o.divergence_matrix([[u] for u in s[:40]], windows=windows, mode=mode, span_normalise=False)
        if mode == "site":
            assert np.array_equal(total_all, want_all)      # integer counts: exact for any number of ranges
            assert np.allclose(total, want, rtol=1e-12, atol=0)
        else:
            assert np.allclose(total, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())
            assert np.allclose(total_all, want_all, rtol=1e-9, atol=1e-9 * np.abs(want_all).max())
    blocks = []
    for rng in ranges:
        ll = LLTreeSequence(sharding.restrict_tables(t, *rng), genome_range=rng)
        blocks.append(ll.genotype_matrix(samples=s[::3]))
        ll.close()
    assert np.array_equal(np.concatenate(blocks, axis=0), o.genotype_matrix(samples=s[::3]))
    # the sharded wrapper itself, one rank: same calls, whole results
    sh = sharding.ShardedTreeSequence(t, windows, 0, 1)
    got = sh.divergence_matrix(windows, sample_sets=flat, sample_set_sizes=sizes, mode="site")
    assert np.allclose(got, o.divergence_matrix(sets, windows=windows, mode="site"), rtol=1e-12, atol=0)
    assert np.array_equal(sh.genotype_matrix(), o.genotype_matrix())


@pytest.mark.skipif(not __import__("oracle.ref", fromlist=["ref"]).available(), reason="oracle/_ref not built")
def test_divmat_8192_samples_against_the_compiled_reference(monkeypatch):
    """The tensor-core contraction at a size that crosses the super-block tile order (32 tile rows of
    256, super-blocks of 12) and runs several waves of CTAs: 8192 samples, biallelic (G G^T through TMA;
    the cp.async kernel as cross-check) and with three alleles per site (one-hot tcgen05 path), exact
    against tsk_treeseq_divergence_matrix (trees.c:8684-8826) compiled from the reference sources."""
    from oracle import ref
    from tskit_b200.lowlevel import LLTreeSequence
    from tskit_b200.sim import add_mutations, wright_fisher
    from tskit_b200.tables import Tables
    t = add_mutations(wright_fisher(8192, 60, 1e5, ncross=1, seed=5), 300, seed=2).ensure_derived()
    s = t.samples
    w = np.array([0.0, 0.4 * t.sequence_length, t.sequence_length])
    want = ref.RefTreeSequence(t).divergence_matrix([[u] for u in s], windows=w, mode="site", span_normalise=False)
    ll = LLTreeSequence(t)
    got = ll.divergence_matrix(w, mode="site", span_normalise=False)
    assert np.array_equal(got, want)
    monkeypatch.setenv("TSKB_MATRIX", "cpasync")
    assert np.array_equal(ll.divergence_matrix(w, mode="site", span_normalise=False), want)
    monkeypatch.delenv("TSKB_MATRIX")
    # sample sets of unequal sizes over the same contraction
    sets = [s[:3000], s[3000:3001], s[4000:8192]]
    sizes = np.array([len(x) for x in sets], dtype=np.uint64)
    got = ll.divergence_matrix(w, sample_sets=np.concatenate(sets).astype(np.int32), sample_set_sizes=sizes, mode="site")
    assert np.allclose(got, ref.RefTreeSequence(t).divergence_matrix(sets, windows=w, mode="site"), rtol=1e-12, atol=0)
    ll.close()
    # three alleles: neighbouring sites merged in pairs, the second mutation of a pair derives "2"
    S = t.num_sites // 2 * 2
    t3 = Tables(t.sequence_length, t.nodes_flags, t.nodes_time, t.edges_left, t.edges_right, t.edges_parent,
                t.edges_child, sites_position=t.sites_position[:S:2],
                mutations_site=(np.arange(S) // 2).astype(np.int32), mutations_node=t.mutations_node[:S],
                mutations_derived_state=np.where(np.arange(S) % 2 == 0, ord("1"), ord("2")).astype(np.int8),
                mutations_derived_state_offset=np.arange(S + 1, dtype=np.uint64),
                edge_insertion_order=t.edge_insertion_order, edge_removal_order=t.edge_removal_order).ensure_derived()
    want3 = ref.RefTreeSequence(t3).divergence_matrix([[u] for u in s], windows=w, mode="site", span_normalise=False)
    ll3 = LLTreeSequence(t3)
    assert np.array_equal(ll3.divergence_matrix(w, mode="site", span_normalise=False), want3)
    ll3.close()
