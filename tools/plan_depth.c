/* plan_depth.c -- analysis tool (not product, not oracle): replays the edge-diff sweep of
 * c/tskit/trees.c:1424-1507 on the host and measures the dependency depth of the
 * count propagation under different level assignments.  Used to choose the propagation
 * schedule of tskit_b200/csrc (see DESIGN.md "Propagation schedule"). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* out[0]=nev out[1]=V out[2]=D_node out[3]=D_entry; hist_entry[lvl] += entries */
int plan_depth(uint64_t N, uint64_t E, double L, const double *left, const double *right,
    const int32_t *parent, const int32_t *child, const int32_t *I, const int32_t *O,
    uint64_t *out, uint64_t *hist_entry, uint64_t hist_len, uint64_t *hist_node,
    uint64_t *hist_chain, uint64_t chain_len)
{
    int32_t *par = malloc(N * sizeof(int32_t));
    uint32_t *cur = calloc(N, sizeof(uint32_t));
    uint32_t *nlevel = calloc(N, sizeof(uint32_t));
    int64_t *stamp = malloc(N * sizeof(int64_t)); uint64_t pairs = 0, nbp = 0; memset(stamp, 0xff, N * sizeof(int64_t));
    uint64_t tj = 0, tk = 0, nev = 0, V = 0;
    uint32_t dmax = 0;
    double t_left = 0;
    memset(par, 0xff, N * sizeof(int32_t));
    /* node levels over the union DAG: edges are sorted by parent time in canonical tables,
     * relax until fixed point */
    int changed = 1;
    while (changed) {
        changed = 0;
        for (uint64_t e = 0; e < E; e++) {
            uint32_t lc = nlevel[child[e]] + 1;
            if (nlevel[parent[e]] < lc) { nlevel[parent[e]] = lc; changed = 1; }
        }
    }
    uint32_t dnode = 0;
    for (uint64_t u = 0; u < N; u++) if (nlevel[u] > dnode) dnode = nlevel[u];
    while (tj < E || t_left < L) {
        nbp++;
        while (tk < E && right[O[tk]] == t_left) {
            int32_t h = O[tk++], c = child[h];
            int32_t u = parent[h];
            uint32_t a = cur[c] + 1;
            uint64_t d = 0;
            par[c] = -1;
            nev++;
            if (stamp[c] != (int64_t) nbp) { stamp[c] = nbp; pairs++; }
            while (u != -1) {
                if (cur[u] < a) cur[u] = a;
                if (cur[u] < hist_len) hist_entry[cur[u]]++;
                if (nlevel[u] < hist_len) hist_node[nlevel[u]]++;
                V++; d++;
                if (stamp[u] != (int64_t) nbp) { stamp[u] = nbp; pairs++; }
                u = par[u];
            }
            if (a > dmax) dmax = a;
            hist_chain[d < chain_len ? d : chain_len - 1]++;
        }
        while (tj < E && left[I[tj]] == t_left) {
            int32_t h = I[tj++], c = child[h];
            int32_t u = parent[h];
            uint32_t a = cur[c] + 1;
            uint64_t d = 0;
            par[c] = u;
            nev++;
            if (stamp[c] != (int64_t) nbp) { stamp[c] = nbp; pairs++; }
            while (u != -1) {
                if (cur[u] < a) cur[u] = a;
                if (cur[u] < hist_len) hist_entry[cur[u]]++;
                if (nlevel[u] < hist_len) hist_node[nlevel[u]]++;
                V++; d++;
                if (stamp[u] != (int64_t) nbp) { stamp[u] = nbp; pairs++; }
                u = par[u];
            }
            if (a > dmax) dmax = a;
            hist_chain[d < chain_len ? d : chain_len - 1]++;
        }
        double t_right = L;
        if (tj < E && left[I[tj]] < t_right) t_right = left[I[tj]];
        if (tk < E && right[O[tk]] < t_right) t_right = right[O[tk]];
        t_left = t_right;
    }
    uint32_t dentry = 0;
    for (uint64_t u = 0; u < N; u++) if (cur[u] > dentry) dentry = cur[u];
    out[4] = pairs; out[5] = nbp;
    out[0] = nev; out[1] = V; out[2] = dnode; out[3] = dentry;
    free(par); free(cur); free(nlevel);
    return 0;
}
