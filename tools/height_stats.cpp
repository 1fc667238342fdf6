// height_stats.cpp -- analysis tool for the "from-scratch piece" schedule: for every (node,
// breakpoint) piece, the number of children of the node in the tree right of the breakpoint and
// the node's height (longest path down to a leaf) there.  Reports the total child references and
// the histogram of heights = the dependency levels that schedule would have.
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>
extern "C" int height_stats(uint64_t N, uint64_t E, double L, const double *left, const double *right,
    const int32_t *parent, const int32_t *child, const int32_t *I, const int32_t *O, uint64_t *out,
    uint64_t *hist, uint64_t hist_len)
{
    std::vector<int32_t> par(N, -1), nch(N, 0), height(N, 0);
    std::vector<std::vector<int32_t>> kids(N);
    std::vector<int64_t> stamp(N, -1);
    std::vector<int32_t> touched;
    uint64_t tj = 0, tk = 0, pieces = 0, refs = 0, nbp = 0, maxh = 0;
    double t_left = 0;
    auto fix_height = [&](int32_t u) {   // recompute heights up the chain from u
        while (u != -1) {
            int32_t h = 0;
            for (int32_t c : kids[u]) h = std::max(h, height[c] + 1);
            height[u] = h;
            u = par[u];
        }
    };
    while (tj < E || t_left < L) {
        nbp++;
        touched.clear();
        auto touch = [&](int32_t u) { if (stamp[u] != (int64_t) nbp) { stamp[u] = nbp; touched.push_back(u); } };
        while (tk < E && right[O[tk]] == t_left) {
            int32_t h = O[tk++], c = child[h], p = parent[h];
            par[c] = -1;
            auto &k = kids[p]; k.erase(std::find(k.begin(), k.end(), c));
            touch(c);
            for (int32_t u = p; u != -1; u = par[u]) touch(u);
            fix_height(p);
        }
        while (tj < E && left[I[tj]] == t_left) {
            int32_t h = I[tj++], c = child[h], p = parent[h];
            par[c] = p;
            kids[p].push_back(c);
            touch(c);
            for (int32_t u = p; u != -1; u = par[u]) touch(u);
            fix_height(p);
        }
        for (int32_t u : touched) {
            pieces++;
            refs += kids[u].size();
            uint64_t h = height[u];
            if (h > maxh) maxh = h;
            if (h < hist_len) hist[h]++;
        }
        double t_right = L;
        if (tj < E && left[I[tj]] < t_right) t_right = left[I[tj]];
        if (tk < E && right[O[tk]] < t_right) t_right = right[O[tk]];
        t_left = t_right;
    }
    out[0] = pieces; out[1] = refs; out[2] = maxh; out[3] = nbp;
    return 0;
}
