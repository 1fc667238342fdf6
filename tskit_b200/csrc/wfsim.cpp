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
// what makes 10^5 samples x 10^7 edges feasible in seconds.  Output tables are in
// tskit's canonical order (edges by (time[parent], parent, child, left), abutting
// edges squashed) together with the edge insertion/removal indexes
// (sort keys of c/tskit/tables.c:11392-11459).
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

namespace {

struct Seg {
    double left, right;
    int32_t node;
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
    Rng(uint64_t seed, uint64_t g, uint64_t i) : key(splitmix(seed ^ splitmix(g * 0x100000001B3ull + i))), ctr(0) {}
    uint64_t next() { return splitmix(key + (ctr++) * 0xD1342543DE82EF95ull); }
    uint64_t below(uint64_t n) { return (uint64_t) (((__uint128_t) next() * n) >> 64); }
};

struct Result {
    std::vector<double> node_time;
    std::vector<uint32_t> node_flags;
    std::vector<Edge> edges;
    std::vector<int32_t> ins, rem;
};

void merge_into_parent(std::vector<Seg> &pieces, int32_t &parent_node, double time,
    Result &R, std::vector<Seg> &out, std::vector<Seg> &X) {
    out.clear();
    if (pieces.size() == 1) {
        out.push_back(pieces[0]);
        return;
    }
    // min-heap on left
    auto cmp = [](const Seg &a, const Seg &b) { return a.left > b.left; };
    std::make_heap(pieces.begin(), pieces.end(), cmp);
    auto push_out = [&](const Seg &a) {
        if (!out.empty() && out.back().node == a.node && out.back().right == a.left) {
            out.back().right = a.right;
        } else {
            out.push_back(a);
        }
    };
    while (!pieces.empty()) {
        double l = pieces.front().left;
        double r = 1e300;
        X.clear();
        while (!pieces.empty() && pieces.front().left == l) {
            std::pop_heap(pieces.begin(), pieces.end(), cmp);
            Seg x = pieces.back();
            pieces.pop_back();
            if (x.right < r) r = x.right;
            X.push_back(x);
        }
        if (!pieces.empty() && pieces.front().left < r) r = pieces.front().left;
        if (X.size() == 1) {
            Seg x = X[0];
            Seg alpha = x;
            if (!pieces.empty() && pieces.front().left < x.right) {
                alpha.right = pieces.front().left;
                x.left = pieces.front().left;
                pieces.push_back(x);
                std::push_heap(pieces.begin(), pieces.end(), cmp);
            }
            push_out(alpha);
        } else {
            if (parent_node < 0) {
                parent_node = (int32_t) R.node_time.size();
                R.node_time.push_back(time);
                R.node_flags.push_back(0);
            }
            for (Seg &x : X) {
                R.edges.push_back(Edge{ l, r, parent_node, x.node });
                if (x.right > r) {
                    x.left = r;
                    pieces.push_back(x);
                    std::push_heap(pieces.begin(), pieces.end(), cmp);
                }
            }
            push_out(Seg{ l, r, parent_node });
        }
    }
}

}  // namespace

extern "C" {

void *tskb_wfsim_run(uint64_t n, uint64_t generations, double L, uint32_t ncross, uint64_t seed) {
    Result *Rp = new Result();
    Result &R = *Rp;
    R.node_time.assign(n, 0.0);
    R.node_flags.assign(n, 1);
    std::vector<std::vector<Seg>> cur(n), nxt(n);
    std::vector<uint32_t> active(n), next_active;
    for (uint64_t i = 0; i < n; i++) {
        cur[i].push_back(Seg{ 0.0, L, (int32_t) i });
        active[i] = (uint32_t) i;
    }
    std::vector<double> bps;
    std::vector<Seg> out, X;
    const uint64_t Lint = (uint64_t) L;
    for (uint64_t g = 1; g <= generations; g++) {
        next_active.clear();
        for (uint32_t idx : active) {
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
                        if (nxt[p].empty()) next_active.push_back(p);
                        nxt[p].push_back(Seg{ s.left, bps[b], s.node });
                        s.left = bps[b];
                        b++;
                    } else {
                        if (nxt[p].empty()) next_active.push_back(p);
                        nxt[p].push_back(s);
                        break;
                    }
                }
            }
            segs.clear();
        }
        // ids must not depend on hash-bucket order: visit parents in id order
        std::sort(next_active.begin(), next_active.end());
        for (uint32_t p : next_active) {
            int32_t parent_node = -1;
            merge_into_parent(nxt[p], parent_node, (double) g, R, out, X);
            cur[p].assign(out.begin(), out.end());
            nxt[p].clear();
        }
        active.swap(next_active);
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
    std::sort(R.ins.begin(), R.ins.end(), [&](int32_t a, int32_t b) {
        if (E[a].left != E[b].left) return E[a].left < E[b].left;
        if (E[a].parent != E[b].parent) return E[a].parent < E[b].parent;
        return E[a].child < E[b].child;
    });
    std::sort(R.rem.begin(), R.rem.end(), [&](int32_t a, int32_t b) {
        if (E[a].right != E[b].right) return E[a].right < E[b].right;
        if (E[a].parent != E[b].parent) return E[a].parent > E[b].parent;
        return E[a].child > E[b].child;
    });
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
