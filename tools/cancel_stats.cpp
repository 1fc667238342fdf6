// cancel_stats.cpp -- analysis tool: how many addends of the node-major plan cancel
// structurally within one breakpoint (removal and re-insertion of the same, unchanged subtree
// below a common ancestor), and how many (node, breakpoint) pieces survive.
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>
extern "C" int cancel_stats(uint64_t N, uint64_t E, double L, const double *left, const double *right,
    const int32_t *parent, const int32_t *child, const int32_t *I, const int32_t *O, uint64_t *out)
{
    std::vector<int32_t> par(N, -1);
    std::vector<int64_t> vis(N, -1), chi(N, -1);
    struct Add { int32_t u; int32_t key; int32_t sign; };
    std::vector<Add> adds;
    struct Ev { int32_t c, p, sign; };
    std::vector<Ev> evs;
    std::vector<std::vector<int32_t>> chains;
    uint64_t tj = 0, tk = 0, V = 0, kept = 0, pieces_all = 0, pieces_kept = 0, nbp = 0, nev = 0;
    double t_left = 0;
    while (tj < E || t_left < L) {
        nbp++;
        evs.clear(); chains.clear();
        while (tk < E && right[O[tk]] == t_left) {
            int32_t h = O[tk++], c = child[h], u = parent[h];
            par[c] = -1;
            evs.push_back({c, parent[h], -1});
            chains.emplace_back();
            while (u != -1) { chains.back().push_back(u); vis[u] = nbp; u = par[u]; }
        }
        while (tj < E && left[I[tj]] == t_left) {
            int32_t h = I[tj++], c = child[h], u = parent[h];
            par[c] = u;
            evs.push_back({c, parent[h], 1});
            chains.emplace_back();
            while (u != -1) { chains.back().push_back(u); vis[u] = nbp; u = par[u]; }
        }
        nev += evs.size();
        adds.clear();
        for (size_t e = 0; e < evs.size(); e++) {
            // the child's state is structurally unchanged at this breakpoint iff it was not visited
            int32_t key = vis[evs[e].c] == (int64_t) nbp ? -(int32_t) e - 2 : evs[e].c;
            chi[evs[e].c] = nbp;
            for (int32_t u : chains[e]) adds.push_back({u, key, evs[e].sign});
        }
        V += adds.size();
        std::sort(adds.begin(), adds.end(), [](const Add &a, const Add &b) {
            return a.u != b.u ? a.u < b.u : a.key < b.key; });
        size_t i = 0;
        int32_t last_u = -1; bool u_kept = false;
        while (i < adds.size()) {
            size_t j = i; int net = 0;
            while (j < adds.size() && adds[j].u == adds[i].u && adds[j].key == adds[i].key) net += adds[j++].sign;
            if (adds[i].u != last_u) {
                if (last_u != -1) { pieces_all++; if (u_kept || chi[last_u] == (int64_t) nbp) pieces_kept++; }
                last_u = adds[i].u; u_kept = false;
            }
            if (net != 0) { kept += (net < 0 ? -net : net); u_kept = true; }
            i = j;
        }
        if (last_u != -1) { pieces_all++; if (u_kept || chi[last_u] == (int64_t) nbp) pieces_kept++; }
        // child-only pieces (nodes that are a child of an event but not visited)
        for (auto &ev : evs) if (vis[ev.c] != (int64_t) nbp && chi[ev.c] == (int64_t) nbp) { chi[ev.c] = -1; pieces_all++; pieces_kept++; }
        double t_right = L;
        if (tj < E && left[I[tj]] < t_right) t_right = left[I[tj]];
        if (tk < E && right[O[tk]] < t_right) t_right = right[O[tk]];
        t_left = t_right;
    }
    out[0] = nev; out[1] = V; out[2] = kept; out[3] = pieces_all; out[4] = pieces_kept; out[5] = nbp;
    return 0;
}
