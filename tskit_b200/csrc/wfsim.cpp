// wfsim.cpp -- seeded synthetic ARG generator (host only; input generation, not on the hot path).
//
// Haploid Wright-Fisher population of constant size n with recombination, the
// model of the reference's forward simulators (c/examples/haploid_wright_fisher.c:12-72,
// python/tests/test_wright_fisher.py:40-206): every individual of generation g draws two
// parents uniformly from generation g+1 and `ncross` crossover positions
// uniform on the integers [1, L-1]; material between crossovers alternates
// between the two parents.  The final generation is the sample.
//
// A forward simulation followed by TableCollection.simplify keeps only the
// ancestry of the samples.  This generator produces that simplified result
// directly by running the same pedigree process backwards in time and carrying
// only ancestral segments (the per-individual segment merge below is the
// overlap-merging step of simplify, Kelleher et al. 2018 algorithm S), which is
// what makes 10^5 samples x 10^7 edges feasible.  Output tables are in
// tskit's canonical order (edges by (time[parent], parent, child, left), abutting
// edges squashed) together with the edge insertion/removal indexes
// (sort keys of c/tskit/tables.c:11392-11459).
//
// Every random draw is a pure function of (seed, generation, individual), and node
// ids are assigned in individual order, so the output is bit-identical for any
// thread count.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <numeric>
#include <thread>
#include <vector>

namespace {

struct Seg {
    double left, right;
    int32_t node;
};

struct Piece {
    uint32_t parent;
    Seg seg;
};

struct Edge {
    double left, right;
    int32_t parent, child;
};

inline uint64_t splitmix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// counter-based stream: draw k of (seed, generation, individual)
struct Rng {
    uint64_t key, ctr;
    Rng(uint64_t seed, uint64_t g, uint64_t i)
        : key(splitmix(seed ^ splitmix(g * 0x100000001B3ull + i))), ctr(0) {}
    uint64_t next() { return splitmix(key + (ctr++) * 0xD1342543DE82EF95ull); }
    uint64_t below(uint64_t n) { return (uint64_t) (((__uint128_t) next() * n) >> 64); }
};

struct Result {
    std::vector<double> node_time;
    std::vector<uint32_t> node_flags;
    std::vector<Edge> edges;
    std::vector<int32_t> ins, rem;
};

constexpr int32_t PENDING = INT32_MIN;  // node id of a coalescence not yet numbered

// Merge the pieces inherited by one parent individual.  Overlaps of >= 2 pieces
// coalesce into the parent's node (id assigned later: PENDING); everything else
// passes through with its own node id.  Returns true if the parent coalesced.
bool merge_into_parent(Seg *pieces, size_t count, std::vector<Edge> &edges,
    std::vector<Seg> &out, std::vector<Seg> &heap, std::vector<Seg> &X) {
    out.clear();
    if (count == 1) {
        out.push_back(pieces[0]);
        return false;
    }
    heap.assign(pieces, pieces + count);
    auto cmp = [](const Seg &a, const Seg &b) { return a.left > b.left; };
    std::make_heap(heap.begin(), heap.end(), cmp);
    bool coalesced = false;
    auto push_out = [&](const Seg &a) {
        if (!out.empty() && out.back().node == a.node && out.back().right == a.left) {
            out.back().right = a.right;
        } else {
            out.push_back(a);
        }
    };
    while (!heap.empty()) {
        double l = heap.front().left;
        double r = 1e300;
        X.clear();
        while (!heap.empty() && heap.front().left == l) {
            std::pop_heap(heap.begin(), heap.end(), cmp);
            Seg x = heap.back();
            heap.pop_back();
            if (x.right < r) r = x.right;
            X.push_back(x);
        }
        if (!heap.empty() && heap.front().left < r) r = heap.front().left;
        if (X.size() == 1) {
            Seg x = X[0];
            Seg alpha = x;
            if (!heap.empty() && heap.front().left < x.right) {
                alpha.right = heap.front().left;
                x.left = heap.front().left;
                heap.push_back(x);
                std::push_heap(heap.begin(), heap.end(), cmp);
            }
            push_out(alpha);
        } else {
            coalesced = true;
            for (Seg &x : X) {
                edges.push_back(Edge{ l, r, PENDING, x.node });
                if (x.right > r) {
                    x.left = r;
                    heap.push_back(x);
                    std::push_heap(heap.begin(), heap.end(), cmp);
                }
            }
            push_out(Seg{ l, r, PENDING });
        }
    }
    return coalesced;
}

// sense-reversing spin barrier for the SPMD generation loop
struct Barrier {
    std::atomic<unsigned> count{0};
    std::atomic<unsigned> sense{0};
    unsigned n;
    explicit Barrier(unsigned n_) : n(n_) {}
    void wait() {
        unsigned s = sense.load(std::memory_order_acquire);
        if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
            count.store(0, std::memory_order_relaxed);
            sense.store(s + 1, std::memory_order_release);
        } else {
            unsigned spins = 0;
            while (sense.load(std::memory_order_acquire) == s) {
                if (++spins > 2000) std::this_thread::yield();
            }
        }
    }
};

}  // namespace

extern "C" {

void *tskb_wfsim_run(uint64_t n, uint64_t generations, double L, uint32_t ncross, uint64_t seed,
    uint32_t num_threads) {
    unsigned nt = num_threads ? num_threads : std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    Result *Rp = new Result();
    Result &R = *Rp;
    R.node_time.assign(n, 0.0);
    R.node_flags.assign(n, 1);
    std::vector<std::vector<Seg>> cur(n);
    std::vector<uint32_t> active(n), next_active;
    for (uint64_t i = 0; i < n; i++) {
        cur[i].push_back(Seg{ 0.0, L, (int32_t) i });
        active[i] = (uint32_t) i;
    }
    const uint64_t Lint = (uint64_t) L;
    std::vector<std::vector<Piece>> tpieces(nt);
    std::vector<std::vector<Edge>> tedges(nt);
    std::vector<std::atomic<uint32_t>> count(n);
    std::vector<uint32_t> start(n + 1), cursor_init(n);
    std::vector<std::atomic<uint32_t>> cursor(n);
    std::vector<Seg> flat;
    std::vector<uint8_t> coalesced(n);
    std::vector<int32_t> newid(n);
    for (uint64_t i = 0; i < n; i++) count[i].store(0, std::memory_order_relaxed);

    if (nt > 1 && n / nt < 2000) nt = (unsigned) std::max<uint64_t>(1, n / 2000);
    tpieces.resize(nt);
    tedges.resize(nt);
    Barrier bar(nt);
    size_t np = 0;
    auto worker = [&](unsigned t) {
        std::vector<double> bps;
        std::vector<Seg> out, heap, X;
        for (uint64_t g = 1; g <= generations; g++) {
            const size_t na = active.size();
            // phase 1: meiosis -- split every lineage's segments between its two parents
            {
                std::vector<Piece> &P = tpieces[t];
                P.clear();
                size_t a0 = na * t / nt, a1 = na * (t + 1) / nt;
                for (size_t a = a0; a < a1; a++) {
                    uint32_t idx = active[a];
                    std::vector<Seg> &segs = cur[idx];
                    Rng rng(seed, g, idx);
                    uint32_t par[2] = { (uint32_t) rng.below(n), (uint32_t) rng.below(n) };
                    bps.clear();
                    for (uint32_t k = 0; k < ncross && Lint > 1; k++) {
                        bps.push_back((double) (1 + rng.below(Lint - 1)));
                    }
                    std::sort(bps.begin(), bps.end());
                    bps.erase(std::unique(bps.begin(), bps.end()), bps.end());
                    size_t b = 0;  // number of crossovers <= current position
                    for (const Seg &s0 : segs) {
                        Seg s = s0;
                        while (b < bps.size() && bps[b] <= s.left) b++;
                        while (true) {
                            uint32_t p = par[b & 1];
                            if (b < bps.size() && bps[b] < s.right) {
                                P.push_back(Piece{ p, Seg{ s.left, bps[b], s.node } });
                                s.left = bps[b];
                                b++;
                            } else {
                                P.push_back(Piece{ p, s });
                                break;
                            }
                        }
                    }
                    segs.clear();
                }
                for (const Piece &pc : P) count[pc.parent].fetch_add(1, std::memory_order_relaxed);
            }
            bar.wait();
            // phase 2: group the pieces by parent (counting sort)
            if (t == 0) {
                next_active.clear();
                uint32_t total = 0;
                for (uint64_t p = 0; p < n; p++) {
                    uint32_t c = count[p].load(std::memory_order_relaxed);
                    start[p] = total;
                    if (c) next_active.push_back((uint32_t) p);
                    total += c;
                    cursor[p].store(start[p], std::memory_order_relaxed);
                    count[p].store(0, std::memory_order_relaxed);
                }
                start[n] = total;
                flat.resize(total);
                np = next_active.size();
            }
            bar.wait();
            for (const Piece &pc : tpieces[t]) {
                flat[cursor[pc.parent].fetch_add(1, std::memory_order_relaxed)] = pc.seg;
            }
            bar.wait();
            // phase 3: merge per parent; coalescences get PENDING node ids
            size_t a0 = np * t / nt, a1 = np * (t + 1) / nt;
            for (size_t a = a0; a < a1; a++) {
                uint32_t p = next_active[a];
                size_t e0 = tedges[t].size();
                bool c = merge_into_parent(flat.data() + start[p], start[p + 1] - start[p],
                    tedges[t], out, heap, X);
                coalesced[p] = c;
                // remember which individual the PENDING ids of these edges belong to
                for (size_t e = e0; e < tedges[t].size(); e++) {
                    tedges[t][e].parent = -(int32_t) p - 2;
                }
                cur[p].assign(out.begin(), out.end());
            }
            bar.wait();
            // phase 4: number the new nodes in individual order (thread-count independent)
            if (t == 0) {
                for (uint32_t p : next_active) {
                    if (coalesced[p]) {
                        newid[p] = (int32_t) R.node_time.size();
                        R.node_time.push_back((double) g);
                        R.node_flags.push_back(0);
                    }
                }
            }
            bar.wait();
            for (size_t a = a0; a < a1; a++) {
                uint32_t p = next_active[a];
                if (coalesced[p]) {
                    for (Seg &s : cur[p]) {
                        if (s.node == PENDING) s.node = newid[p];
                    }
                }
            }
            for (Edge &e : tedges[t]) {
                if (e.parent < -1) e.parent = newid[(uint32_t) (-(e.parent + 2))];
            }
            bar.wait();
            if (t == 0) active.swap(next_active);
            bar.wait();
        }
    };
    {
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nt; t++) th.emplace_back(worker, t);
        worker(0);
        for (auto &x : th) x.join();
    }
    for (unsigned t = 0; t < nt; t++) {
        R.edges.insert(R.edges.end(), tedges[t].begin(), tedges[t].end());
        std::vector<Edge>().swap(tedges[t]);
    }
    // squash abutting edges and put them in canonical order
    std::vector<Edge> &E = R.edges;
    std::sort(E.begin(), E.end(), [](const Edge &a, const Edge &b) {
        if (a.parent != b.parent) return a.parent < b.parent;
        if (a.child != b.child) return a.child < b.child;
        return a.left < b.left;
    });
    size_t m = 0;
    for (size_t j = 0; j < E.size(); j++) {
        if (m > 0 && E[m - 1].parent == E[j].parent && E[m - 1].child == E[j].child
            && E[m - 1].right == E[j].left) {
            E[m - 1].right = E[j].right;
        } else {
            E[m++] = E[j];
        }
    }
    E.resize(m);
    // edge indexes; node time is non-decreasing in node id, so (time, id) order is id order
    R.ins.resize(m);
    R.rem.resize(m);
    std::iota(R.ins.begin(), R.ins.end(), 0);
    std::iota(R.rem.begin(), R.rem.end(), 0);
    std::thread ti([&]() {
        std::sort(R.ins.begin(), R.ins.end(), [&](int32_t a, int32_t b) {
            if (E[a].left != E[b].left) return E[a].left < E[b].left;
            if (E[a].parent != E[b].parent) return E[a].parent < E[b].parent;
            return E[a].child < E[b].child;
        });
    });
    std::sort(R.rem.begin(), R.rem.end(), [&](int32_t a, int32_t b) {
        if (E[a].right != E[b].right) return E[a].right < E[b].right;
        if (E[a].parent != E[b].parent) return E[a].parent > E[b].parent;
        return E[a].child > E[b].child;
    });
    ti.join();
    return Rp;
}

void tskb_wfsim_sizes(const void *h, uint64_t *num_nodes, uint64_t *num_edges) {
    const Result *R = (const Result *) h;
    *num_nodes = R->node_time.size();
    *num_edges = R->edges.size();
}

void tskb_wfsim_copy(const void *h, uint32_t *node_flags, double *node_time, double *left,
    double *right, int32_t *parent, int32_t *child, int32_t *insertion, int32_t *removal) {
    const Result *R = (const Result *) h;
    memcpy(node_flags, R->node_flags.data(), R->node_flags.size() * sizeof(uint32_t));
    memcpy(node_time, R->node_time.data(), R->node_time.size() * sizeof(double));
    for (size_t j = 0; j < R->edges.size(); j++) {
        left[j] = R->edges[j].left;
        right[j] = R->edges[j].right;
        parent[j] = R->edges[j].parent;
        child[j] = R->edges[j].child;
    }
    memcpy(insertion, R->ins.data(), R->ins.size() * sizeof(int32_t));
    memcpy(removal, R->rem.data(), R->rem.size() * sizeof(int32_t));
}

void tskb_wfsim_free(void *h) { delete (Result *) h; }

}  // extern "C"
