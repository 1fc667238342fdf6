"""Multi-GPU host logic on CPU: world_size-2 gloo processes, each computing its genome range with
an engine stand-in built on the oracle (test infrastructure), combined by the product's
sharding module; the result must equal the whole-genome oracle."""
import os

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from tskit_b200 import sharding
from tskit_b200.tables import Tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "tests", "data", "wf_200_500_100000.npz")


class OracleRangeEngine:
    """Statistic of the trees/sites inside [lo, hi) only, from the whole-genome oracle: refine
    the windows with the cuts, evaluate un-normalised, add the sub-windows inside the range."""

    def __init__(self, tables, rng):
        from oracle import port
        self.o = port.Oracle(tables)
        self.lo, self.hi = rng

    def _run(self, name, sets, indexes, windows, mode, **kw):
        w = np.asarray(windows, dtype=np.float64)
        fine = np.unique(np.concatenate([w, [self.lo, self.hi]]))
        r = self.o.stat(name, sets, indexes, windows=fine, mode=mode, span_normalise=False, **kw)
        out = np.zeros((len(w) - 1, r.shape[1]))
        mid = 0.5 * (fine[:-1] + fine[1:])
        inside = (mid >= self.lo) & (mid < self.hi)
        owner = np.searchsorted(w, mid, side="right") - 1
        np.add.at(out, owner[inside], r[inside])
        return out

    def diversity(self, sets, windows, mode, span_normalise):
        assert span_normalise is False
        return self._run("diversity", sets, None, windows, mode)

    def divergence(self, sets, indexes, windows, mode, span_normalise):
        assert span_normalise is False
        return self._run("divergence", sets, indexes, windows, mode)

    def genetic_relatedness_vector(self, weights, windows, mode, span_normalise, centre, nodes):
        assert span_normalise is False and mode == "branch"
        w = np.asarray(windows, dtype=np.float64)
        fine = np.unique(np.concatenate([w, [self.lo, self.hi]]))
        r = self.o.genetic_relatedness_vector(weights, windows=fine, nodes=nodes, centre=centre,
                                              span_normalise=False)
        out = np.zeros((len(w) - 1,) + r.shape[1:])
        mid = 0.5 * (fine[:-1] + fine[1:])
        inside = (mid >= self.lo) & (mid < self.hi)
        owner = np.searchsorted(w, mid, side="right") - 1
        np.add.at(out, owner[inside], r[inside])
        return out


    def divergence_matrix(self, windows, sample_sets, sample_set_sizes, mode, span_normalise):
        assert span_normalise is False
        w = np.asarray(windows, dtype=np.float64)
        fine = np.unique(np.concatenate([w, [self.lo, self.hi]]))
        fine = fine[(fine >= w[0]) & (fine <= w[-1])]
        off = np.concatenate([[0], np.cumsum(sample_set_sizes)]).astype(int)
        sets = [sample_sets[off[i]:off[i + 1]] for i in range(len(sample_set_sizes))]
        r = self.o.divergence_matrix(sets, windows=fine, mode=mode, span_normalise=False)
        out = np.zeros((len(w) - 1,) + r.shape[1:])
        mid = 0.5 * (fine[:-1] + fine[1:])
        inside = (mid >= self.lo) & (mid < self.hi)
        owner = np.searchsorted(w, mid, side="right") - 1
        np.add.at(out, owner[inside], r[inside])
        return out

    def genotype_matrix(self, samples=None, isolated_as_missing=True):
        return self.o.genotype_matrix(samples=samples, isolated_as_missing=isolated_as_missing).astype(np.int8)


def _matrix_worker(rank, world, port_no, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t = Tables.load(DATA).ensure_derived()
    L = t.sequence_length
    windows = np.array([0.0, 0.3 * L, L])
    sh = sharding.ShardedTreeSequence(t, np.linspace(0, L, 21), rank, world, engine_factory=OracleRangeEngine)
    s = t.samples
    sets = [s[:12], s[12:13], s[20:45]]
    sizes = np.array([len(x) for x in sets], dtype=np.uint64)
    flat = np.concatenate(sets).astype(np.int32)
    out = {}
    for mode in ("site", "branch"):
        out[mode] = sh.divergence_matrix(windows, sample_sets=flat, sample_set_sizes=sizes, mode=mode)
    out["genotypes"] = sh.genotype_matrix(samples=s[::4])
    out["block"] = sh.genotype_matrix(samples=s[::4], gather=False)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matrices_and_decode():
    """site / branch divergence matrix summed over genome ranges, genotype decode
    sharded by site -- the product's sharding module over gloo, engines standing in from the oracle."""
    from oracle import port
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 29500 + (os.getpid() + 977) % 2000
    procs = [ctx.Process(target=_matrix_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    t = Tables.load(DATA).ensure_derived()
    o = port.Oracle(t)
    L = t.sequence_length
    windows = np.array([0.0, 0.3 * L, L])
    s = t.samples
    sets = [s[:12], s[12:13], s[20:45]]
    G = o.genotype_matrix(samples=s[::4])
    firsts = {}
    for rank, out in got:
        want = o.divergence_matrix(sets, windows=windows, mode="site")
        assert np.allclose(out["site"], want, rtol=1e-12, atol=0)
        want = o.divergence_matrix(sets, windows=windows, mode="branch")
        assert np.allclose(out["branch"], want, rtol=1e-9, atol=1e-9 * np.abs(want).max())
        assert np.array_equal(out["genotypes"], G)
        first, block = out["block"]
        assert np.array_equal(block, G[first:first + len(block)])
        firsts[rank] = (first, len(block))
    assert firsts[0][0] == 0 and firsts[1][0] == firsts[0][1] and sum(v[1] for v in firsts.values()) == len(G)


def _worker(rank, world, port_no, W, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t = Tables.load(DATA).ensure_derived()
    windows = np.linspace(0, t.sequence_length, W + 1)
    sh = sharding.ShardedTreeSequence(t, windows, rank, world, engine_factory=OracleRangeEngine)
    s = t.samples
    sets = [s[:90], s[90:]]
    out = {}
    for mode in ("branch", "site"):
        out[("diversity", mode)] = sh.stat("diversity", sets, windows=windows, mode=mode)
        out[("divergence", mode)] = sh.stat("divergence", sets, [[0, 1]], windows=windows, mode=mode)
    wt = np.random.default_rng(3).normal(size=(len(s), 2))
    for centre in (True, False):
        out[("relvec", centre)] = sh.stat("genetic_relatedness_vector", wt, windows=windows, mode="branch",
                                          centre=centre, nodes=s)
    q.put((rank, sh.ranges, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("W", [1, 7])
def test_two_rank_gloo_matches_whole_genome(W):
    from oracle import port
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 29500 + (os.getpid() + W) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, W, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    t = Tables.load(DATA).ensure_derived()
    o = port.Oracle(t)
    windows = np.linspace(0, t.sequence_length, W + 1)
    s = t.samples
    sets = [s[:90], s[90:]]
    for rank, ranges, out in got:
        assert len(ranges) == 2 and ranges[0][1] == ranges[1][0]
        if W >= 2:
            assert ranges[0][1] in windows  # cut snapped to a window edge
        for mode in ("branch", "site"):
            want = o.stat("diversity", sets, windows=windows, mode=mode)
            assert np.allclose(out[("diversity", mode)], want, rtol=1e-11)
            want = o.stat("divergence", sets, [[0, 1]], windows=windows, mode=mode)
            assert np.allclose(out[("divergence", mode)], want, rtol=1e-11)
        wt = np.random.default_rng(3).normal(size=(len(s), 2))
        for centre in (True, False):
            want = o.genetic_relatedness_vector(wt, windows=windows, centre=centre)
            got_v = out[("relvec", centre)]
            assert got_v.shape == want.shape
            assert np.allclose(got_v, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())


def test_plan_shards_balanced_and_covering(wf_1k):
    pos = sharding.edge_diff_positions(wf_1k)
    L = wf_1k.sequence_length
    for world, W in ((8, 1000), (4, 2), (2, 1)):
        windows = np.linspace(0, L, W + 1)
        r = sharding.plan_shards(wf_1k, windows, world)
        assert len(r) == world and r[0][0] == 0.0 and r[-1][1] == L
        assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
        counts = [np.count_nonzero((pos >= lo) & (pos < hi)) for lo, hi in r]
        assert max(counts) < 1.5 * len(pos) / world


def test_restricted_tables_of_a_repeated_genome_sum_to_whole(wf_small):
    """The pieces bench.py --gpus N is built from: repeat_genome (weak-scaling workload),
    plan_shards on the edge indexes, restrict_tables per range.  Every range evaluated by the oracle
    on its restricted tables; the sum over ranges is the oracle on the whole tables, and every copy
    of the genome repeats the statistic of the original."""
    from oracle import port
    from tskit_b200.sim import repeat_genome
    base = wf_small
    copies, W = 3, 5
    t = repeat_genome(base, copies)
    assert t.num_edges == copies * base.num_edges and t.num_samples == base.num_samples
    windows = np.linspace(0, t.sequence_length, copies * W + 1)
    ranges = sharding.plan_shards(t, windows, 4)
    s = t.samples
    sets = [s[: len(s) // 3], s[len(s) // 3:]]
    whole = port.Oracle(t)
    for mode in ("branch", "site"):
        want = whole.stat("divergence", sets, [[0, 1]], windows=windows, mode=mode, span_normalise=False)
        total = np.zeros_like(want)
        for rng in ranges:
            local = sharding.restrict_tables(t, *rng)
            assert local.num_edges < t.num_edges and local.num_nodes == t.num_nodes
            total += OracleRangeEngine(local, rng).divergence(sets, [[0, 1]], windows, mode, False)
        assert np.allclose(total, want, rtol=1e-11, atol=0)
        one = port.Oracle(base).stat("divergence", sets, [[0, 1]],
                                     windows=np.linspace(0, base.sequence_length, W + 1), mode=mode,
                                     span_normalise=False)
        assert np.allclose(want.reshape(copies, W), one.reshape(1, W), rtol=1e-11, atol=0)


def _fallback_worker(rank, world, port_no, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t = Tables.load(DATA).ensure_derived()
    L = t.sequence_length
    windows = np.linspace(0, L, 6)
    sh = sharding.ShardedTreeSequence(t, windows, rank, world, engine_factory=OracleRangeEngine)
    ex = sh.use_peer_exchange(64)   # no CUDA device here: every rank must agree on "no exchange"
    s = t.samples
    got = sh.stat("diversity", [s[:30], s[30:]], windows=windows, mode="branch")
    q.put((rank, ex is None, sh.exchange_note, got))
    dist.barrier()
    dist.destroy_process_group()


def test_peer_exchange_falls_back_collectively_without_a_device():
    """use_peer_exchange where the receive slab cannot be created (this container has no GPU): every
    rank gets None after the same collectives (no rank left waiting in one) and the statistics go on
    through the all_reduce."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a host without a CUDA device")
    from oracle import port
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 29500 + (os.getpid() + 1483) % 2000
    procs = [ctx.Process(target=_fallback_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    t = Tables.load(DATA).ensure_derived()
    s = t.samples
    want = port.Oracle(t).stat("diversity", [s[:30], s[30:]], windows=np.linspace(0, t.sequence_length, 6),
                               mode="branch")
    for rank, is_none, note, got in res:
        assert is_none and note
        assert np.allclose(got, want, rtol=1e-9)
