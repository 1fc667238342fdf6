// noop_stats.cpp -- analysis tool: how many (node, breakpoint) pieces are structural no-ops?
// Sweeps the trees left to right keeping, for every node, a version id of its subtree (a hash of
// its children's version ids; a sample's own id seeds it).  At every breakpoint each touched node
// gets a piece; the piece is a "state no-op" if the node's version equals the one of its previous
// piece (same descendant samples for ANY weights), and a "full no-op" if its parent is unchanged too.
#include <stdint.h>
#include <algorithm>
#include <vector>
static inline uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
extern "C" int noop_stats(uint64_t N, uint64_t E, double L, const double *left, const double *right,
    const int32_t *parent, const int32_t *child, const int32_t *I, const int32_t *O, const double *time,
    const uint32_t *flags, uint64_t *out)
{
    std::vector<int32_t> par(N, -1), last_par(N, -2);
    std::vector<uint64_t> own(N), ver(N), last_ver(N, ~0ull), kidsum(N, 0);
    for (uint64_t u = 0; u < N; u++) { own[u] = (flags[u] & 1) ? mix(u + 1) : 0; ver[u] = own[u]; }
    std::vector<int64_t> stamp(N, -1);
    std::vector<int32_t> touched;
    uint64_t tj = 0, tk = 0, pieces = 0, state_noop = 0, full_noop = 0, nbp = 0, root_pieces = 0, first = 0;
    double t_left = 0;
    while (tj < E || t_left < L) {
        nbp++;
        touched.clear();
        auto touch = [&](int32_t u) { if (stamp[u] != (int64_t) nbp) { stamp[u] = nbp; touched.push_back(u); } };
        // version of u = own + sum over children of mix(version(child)): update along the path
        auto bump = [&](int32_t p, uint64_t delta) {
            for (int32_t u = p; u != -1; u = par[u]) {
                touch(u);
                uint64_t oldc = mix(ver[u]);
                kidsum[u] += delta;
                ver[u] = own[u] + kidsum[u];
                delta = mix(ver[u]) - oldc;
            }
        };
        while (tk < E && right[O[tk]] == t_left) {
            int32_t h = O[tk++], c = child[h], p = parent[h];
            touch(c);
            par[c] = -1;
            bump(p, 0 - mix(ver[c]));
        }
        while (tj < E && left[I[tj]] == t_left) {
            int32_t h = I[tj++], c = child[h], p = parent[h];
            touch(c);
            par[c] = p;
            bump(p, mix(ver[c]));
        }
        for (int32_t u : touched) {
            pieces++;
            if (par[u] == -1) root_pieces++;
            if (last_par[u] == -2) first++;
            else if (ver[u] == last_ver[u]) { state_noop++; if (par[u] == last_par[u]) full_noop++; }
            last_ver[u] = ver[u];
            last_par[u] = par[u];
        }
        double t_right = L;
        if (tj < E && left[I[tj]] < t_right) t_right = left[I[tj]];
        if (tk < E && right[O[tk]] < t_right) t_right = right[O[tk]];
        t_left = t_right;
    }
    out[0] = pieces; out[1] = state_noop; out[2] = full_noop; out[3] = nbp; out[4] = root_pieces; out[5] = first;
    return 0;
}
