// src_stats.cpp -- analysis tool: where do the addend sources of the plan point?
// Classifies every visit by the state of the diff's child: a node that is never a parent
// (state == own weight forever), a node whose state has not changed yet, or a general piece;
// and measures reuse distance proxies (distinct canonical sources).
#include <stdint.h>
#include <vector>
extern "C" int src_stats(uint64_t N, uint64_t E, double L, const double *left, const double *right,
    const int32_t *parent, const int32_t *child, const int32_t *I, const int32_t *O, uint64_t *out)
{
    std::vector<int32_t> par(N, -1);
    std::vector<uint8_t> is_parent(N, 0), touched(N, 0);
    std::vector<uint32_t> nvis(N, 0);
    for (uint64_t e = 0; e < E; e++) is_parent[parent[e]] = 1;
    uint64_t tj = 0, tk = 0, V = 0, v_leaf = 0, v_untouched = 0, ev_leaf = 0, nev = 0;
    double t_left = 0;
    auto walk = [&](int32_t c, int32_t u) {
        uint64_t d = 0;
        while (u != -1) { d++; touched[u] = 1; nvis[u]++; u = par[u]; }
        V += d; nev++;
        if (!is_parent[c]) { v_leaf += d; ev_leaf++; }
        else if (!touched[c]) v_untouched += d;
    };
    while (tj < E || t_left < L) {
        while (tk < E && right[O[tk]] == t_left) { int32_t h = O[tk++]; par[child[h]] = -1; walk(child[h], parent[h]); }
        while (tj < E && left[I[tj]] == t_left) { int32_t h = I[tj++]; par[child[h]] = parent[h]; walk(child[h], parent[h]); }
        double t_right = L;
        if (tj < E && left[I[tj]] < t_right) t_right = left[I[tj]];
        if (tk < E && right[O[tk]] < t_right) t_right = right[O[tk]];
        t_left = t_right;
    }
    // list-length distribution: visits in lists longer than 1e3, 1e4, 1e5
    uint64_t l3 = 0, l4 = 0, l5 = 0, n3 = 0, n4 = 0, n5 = 0, nleaf = 0;
    for (uint64_t u = 0; u < N; u++) {
        if (!is_parent[u]) nleaf++;
        if (nvis[u] > 1000) { l3 += nvis[u]; n3++; }
        if (nvis[u] > 10000) { l4 += nvis[u]; n4++; }
        if (nvis[u] > 100000) { l5 += nvis[u]; n5++; }
    }
    uint64_t r[] = {nev, V, v_leaf, v_untouched, ev_leaf, nleaf, l3, n3, l4, n4, l5, n5};
    for (int i = 0; i < 12; i++) out[i] = r[i];
    return 0;
}
