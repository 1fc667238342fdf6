/*
 * stats_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded restatement of the reference's general-statistics
 * algorithm, written from the reference's description of what it computes:
 * one left-to-right sweep over edge diffs with a root-ward walk per diff.
 * Each function cites the reference file:line it follows (paths relative to
 * the tskit repository).  The oracle is pinned against the reference's golden
 * vectors and against oracle/_ref (the compiled reference itself) by
 * tests/test_oracle.py.  Nothing under tskit_b200/ may link or call this file.
 *
 * Build: gcc -O2 -std=c99 -shared -fPIC stats_oracle.c -lm -o _build/liboracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_NULL (-1)
#define ORC_STAT_SITE (1u << 0)
#define ORC_STAT_BRANCH (1u << 1)
#define ORC_STAT_NODE (1u << 2)
#define ORC_STAT_POLARISED (1u << 10)
#define ORC_STAT_SPAN_NORMALISE (1u << 11)
#define ORC_STAT_ALLOW_TIME_UNCALIBRATED (1u << 12)
#define ORC_STAT_NONCENTRED (1u << 14)
#define ORC_ISOLATED_NOT_MISSING (1u << 1)

/* error codes: c/tskit/core.h:259-698 */
#define ORC_ERR_NO_MEMORY (-2)
#define ORC_ERR_NODE_OUT_OF_BOUNDS (-202)
#define ORC_ERR_DUPLICATE_SAMPLE (-600)
#define ORC_ERR_BAD_SAMPLES (-601)
#define ORC_ERR_BAD_NUM_WINDOWS (-900)
#define ORC_ERR_BAD_WINDOWS (-901)
#define ORC_ERR_MULTIPLE_STAT_MODES (-902)
#define ORC_ERR_BAD_STATE_DIMS (-903)
#define ORC_ERR_BAD_RESULT_DIMS (-904)
#define ORC_ERR_INSUFFICIENT_SAMPLE_SETS (-905)
#define ORC_ERR_INSUFFICIENT_INDEX_TUPLES (-906)
#define ORC_ERR_BAD_SAMPLE_SET_INDEX (-907)
#define ORC_ERR_EMPTY_SAMPLE_SET (-908)
#define ORC_ERR_UNSUPPORTED_STAT_MODE (-909)
#define ORC_ERR_TIME_UNCALIBRATED (-910)
#define ORC_ERR_STAT_POLARISED_UNSUPPORTED (-911)

typedef struct {
    double sequence_length;
    int32_t time_uncalibrated;
    uint64_t num_nodes;
    const uint32_t *node_flags;
    const double *node_time;
    uint64_t num_edges;
    const double *edge_left;
    const double *edge_right;
    const int32_t *edge_parent;
    const int32_t *edge_child;
    const int32_t *edge_insertion_order;
    const int32_t *edge_removal_order;
    uint64_t num_sites;
    const double *site_position;
    const char *site_ancestral_state;
    const uint64_t *site_ancestral_state_offset;
    uint64_t num_mutations;
    const int32_t *mutation_site;
    const int32_t *mutation_node;
    const int32_t *mutation_parent;
    const char *mutation_derived_state;
    const uint64_t *mutation_derived_state_offset;
} orc_tables_t;

typedef int orc_summary_func(uint64_t state_dim, const double *state, uint64_t result_dim,
    double *result, void *params);

/* ------------------------------------------------------------------ sweep
 * Edge-diff iteration: at each breakpoint first every edge ending there (in
 * removal-index order), then every edge starting there (insertion-index
 * order); c/tskit/trees.c:1424-1482. */
typedef struct {
    const orc_tables_t *t;
    int64_t in, out; /* cursors into the insertion / removal indexes */
    double left, right;
} sweep_t;

static void
sweep_init(sweep_t *s, const orc_tables_t *t)
{
    s->t = t;
    s->in = 0;
    s->out = 0;
    s->left = 0;
    s->right = 0;
}

static int
sweep_more(const sweep_t *s)
{
    return s->in < (int64_t) s->t->num_edges || s->left < s->t->sequence_length;
}

/* next edge leaving at the current left coordinate, or -1 */
static int32_t
sweep_next_out(sweep_t *s)
{
    const orc_tables_t *t = s->t;
    if (s->out < (int64_t) t->num_edges
        && t->edge_right[t->edge_removal_order[s->out]] == s->left) {
        return t->edge_removal_order[s->out++];
    }
    return ORC_NULL;
}

static int32_t
sweep_next_in(sweep_t *s)
{
    const orc_tables_t *t = s->t;
    if (s->in < (int64_t) t->num_edges
        && t->edge_left[t->edge_insertion_order[s->in]] == s->left) {
        return t->edge_insertion_order[s->in++];
    }
    return ORC_NULL;
}

/* right end of the tree just built */
static void
sweep_close(sweep_t *s)
{
    const orc_tables_t *t = s->t;
    double r = t->sequence_length;
    if (s->in < (int64_t) t->num_edges) {
        double x = t->edge_left[t->edge_insertion_order[s->in]];
        r = x < r ? x : r;
    }
    if (s->out < (int64_t) t->num_edges) {
        double x = t->edge_right[t->edge_removal_order[s->out]];
        r = x < r ? x : r;
    }
    s->right = r;
}

/* ------------------------------------------------------------ validation */

/* c/tskit/trees.c:1244-1286 */
static int
check_windows(const orc_tables_t *t, uint64_t W, const double *windows, int full_span)
{
    uint64_t j;
    if (W < 1) {
        return ORC_ERR_BAD_NUM_WINDOWS;
    }
    if (full_span) {
        if (windows[0] != 0 || windows[W] != t->sequence_length) {
            return ORC_ERR_BAD_WINDOWS;
        }
    } else if (windows[0] < 0 || windows[W] > t->sequence_length) {
        return ORC_ERR_BAD_WINDOWS;
    }
    for (j = 0; j < W; j++) {
        if (windows[j] >= windows[j + 1]) {
            return ORC_ERR_BAD_WINDOWS;
        }
    }
    return 0;
}

/* sample index map of init_nodes, c/tskit/trees.c:404-453 */
static int32_t *
sample_index_map(const orc_tables_t *t, uint64_t *num_samples)
{
    int32_t *map = malloc((t->num_nodes + 1) * sizeof(*map));
    uint64_t u, n = 0;
    if (map != NULL) {
        for (u = 0; u < t->num_nodes; u++) {
            map[u] = (t->node_flags[u] & 1u) ? (int32_t) n++ : ORC_NULL;
        }
        *num_samples = n;
    }
    return map;
}

/* ------------------------------------------------------- general statistic */

typedef struct {
    orc_summary_func *f;
    void *f_params;
    uint64_t K;
    const double *total;
    double *tmp_state;
    double *tmp_result;
} unpolarised_t;

/* f(x) + f(total - x): c/tskit/trees.c:1944-1972 */
static int
unpolarised(uint64_t K, const double *x, uint64_t M, double *result, void *params)
{
    unpolarised_t *u = params;
    uint64_t k, m;
    int ret = u->f(K, x, M, result, u->f_params);
    if (ret != 0) {
        return ret;
    }
    for (k = 0; k < K; k++) {
        u->tmp_state[k] = u->total[k] - x[k];
    }
    ret = u->f(K, u->tmp_state, M, u->tmp_result, u->f_params);
    for (m = 0; m < M; m++) {
        result[m] += u->tmp_result[m];
    }
    return ret;
}

/* c/tskit/trees.c:1352-1523 */
static int
branch_stat(const orc_tables_t *t, uint64_t n, const int32_t *samples, uint64_t K,
    const double *weights, uint64_t M, orc_summary_func *f, void *fp, uint64_t W,
    const double *windows, double *result)
{
    int ret = 0;
    const uint64_t N = t->num_nodes;
    int32_t *parent = malloc((N + 1) * sizeof(*parent));
    double *blen = calloc(N + 1, sizeof(*blen));
    double *state = calloc((N + 1) * K, sizeof(*state));
    double *summary = calloc((N + 1) * M, sizeof(*summary));
    double *acc = calloc(M, sizeof(*acc));
    double *zero = calloc(K, sizeof(*zero));
    uint64_t j, m, k, w = 0;
    int32_t e, u, c;
    sweep_t s;

    if (!parent || !blen || !state || !summary || !acc || !zero) {
        ret = ORC_ERR_NO_MEMORY;
        goto out;
    }
    memset(parent, 0xff, (N + 1) * sizeof(*parent));
    /* every node starts with f(0); samples with f(their weight) (1396-1415) */
    ret = f(K, zero, M, summary, fp);
    if (ret != 0) {
        goto out;
    }
    for (j = 1; j < N; j++) {
        memcpy(summary + j * M, summary, M * sizeof(double));
    }
    for (j = 0; j < n; j++) {
        u = samples[j];
        memcpy(state + (uint64_t) u * K, weights + j * K, K * sizeof(double));
        ret = f(K, state + (uint64_t) u * K, M, summary + (uint64_t) u * M, fp);
        if (ret != 0) {
            goto out;
        }
    }
    memset(result, 0, W * M * sizeof(double));

#define ACC(node, sgn)                                                                  \
    for (m = 0; m < M; m++) {                                                           \
        acc[m] += (sgn) *blen[node] * summary[(uint64_t)(node) *M + m];                 \
    }
    sweep_init(&s, t);
    while (sweep_more(&s)) {
        while ((e = sweep_next_out(&s)) != ORC_NULL) {
            c = t->edge_child[e];
            ACC(c, -1.0);
            parent[c] = ORC_NULL;
            blen[c] = 0;
            for (u = t->edge_parent[e]; u != ORC_NULL; u = parent[u]) {
                ACC(u, -1.0);
                for (k = 0; k < K; k++) {
                    state[(uint64_t) u * K + k] -= state[(uint64_t) c * K + k];
                }
                ret = f(K, state + (uint64_t) u * K, M, summary + (uint64_t) u * M, fp);
                if (ret != 0) {
                    goto out;
                }
                ACC(u, +1.0);
            }
        }
        while ((e = sweep_next_in(&s)) != ORC_NULL) {
            c = t->edge_child[e];
            u = t->edge_parent[e];
            parent[c] = u;
            blen[c] = t->node_time[u] - t->node_time[c];
            ACC(c, +1.0);
            for (; u != ORC_NULL; u = parent[u]) {
                ACC(u, -1.0);
                for (k = 0; k < K; k++) {
                    state[(uint64_t) u * K + k] += state[(uint64_t) c * K + k];
                }
                ret = f(K, state + (uint64_t) u * K, M, summary + (uint64_t) u * M, fp);
                if (ret != 0) {
                    goto out;
                }
                ACC(u, +1.0);
            }
        }
        sweep_close(&s);
        /* integrate the running sum over the overlap with each window (1484-1504) */
        while (w < W && windows[w] < s.right) {
            double lo = s.left > windows[w] ? s.left : windows[w];
            double hi = s.right < windows[w + 1] ? s.right : windows[w + 1];
            for (m = 0; m < M; m++) {
                result[w * M + m] += acc[m] * (hi - lo);
            }
            if (windows[w + 1] <= s.right) {
                w++;
            } else {
                break;
            }
        }
        s.left = s.right;
    }
#undef ACC
out:
    free(parent);
    free(blen);
    free(state);
    free(summary);
    free(acc);
    free(zero);
    return ret;
}

/* allele states of one site: c/tskit/trees.c:1525-1612 */
static int
site_result(const orc_tables_t *t, uint64_t site, uint64_t m0, uint64_t m1, const double *state,
    uint64_t K, const double *total, uint64_t M, orc_summary_func *f, void *fp, int polarised,
    double *out, double *tmp)
{
    int ret = 0;
    uint64_t max_alleles = m1 - m0 + 1, num_alleles = 1, a, k, m, j;
    const char **allele = malloc(max_alleles * sizeof(*allele));
    uint64_t *allele_len = malloc(max_alleles * sizeof(*allele_len));
    double *astate = calloc(max_alleles * K, sizeof(*astate));
    const char *alt;
    uint64_t alt_len;

    if (!allele || !allele_len || !astate) {
        ret = ORC_ERR_NO_MEMORY;
        goto out;
    }
    allele[0] = t->site_ancestral_state + t->site_ancestral_state_offset[site];
    allele_len[0]
        = t->site_ancestral_state_offset[site + 1] - t->site_ancestral_state_offset[site];
    memcpy(astate, total, K * sizeof(double));
    for (j = m0; j < m1; j++) {
        const char *der = t->mutation_derived_state + t->mutation_derived_state_offset[j];
        uint64_t der_len
            = t->mutation_derived_state_offset[j + 1] - t->mutation_derived_state_offset[j];
        const double *x = state + (uint64_t) t->mutation_node[j] * K;
        for (a = 0; a < num_alleles; a++) {
            if (allele_len[a] == der_len && memcmp(allele[a], der, der_len) == 0) {
                break;
            }
        }
        if (a == num_alleles) {
            allele[a] = der;
            allele_len[a] = der_len;
            num_alleles++;
        }
        for (k = 0; k < K; k++) {
            astate[a * K + k] += x[k];
        }
        alt = allele[0];
        alt_len = allele_len[0];
        if (t->mutation_parent[j] != ORC_NULL) {
            int32_t pm = t->mutation_parent[j];
            alt = t->mutation_derived_state + t->mutation_derived_state_offset[pm];
            alt_len = t->mutation_derived_state_offset[pm + 1]
                      - t->mutation_derived_state_offset[pm];
        }
        for (a = 0; a < num_alleles; a++) {
            if (allele_len[a] == alt_len && memcmp(allele[a], alt, alt_len) == 0) {
                break;
            }
        }
        for (k = 0; k < K; k++) {
            astate[a * K + k] -= x[k];
        }
    }
    /* sum f over alleles, skipping the ancestral one when polarised (1614-1652) */
    memset(out, 0, M * sizeof(double));
    for (a = polarised ? 1 : 0; a < num_alleles; a++) {
        ret = f(K, astate + a * K, M, tmp, fp);
        if (ret != 0) {
            goto out;
        }
        for (m = 0; m < M; m++) {
            out[m] += tmp[m];
        }
    }
out:
    free(allele);
    free(allele_len);
    free(astate);
    return ret;
}

/* c/tskit/trees.c:1654-1776 */
static int
site_stat(const orc_tables_t *t, uint64_t n, const int32_t *samples, uint64_t K,
    const double *weights, uint64_t M, orc_summary_func *f, void *fp, uint64_t W,
    const double *windows, int polarised, double *result)
{
    int ret = 0;
    const uint64_t N = t->num_nodes;
    int32_t *parent = malloc((N + 1) * sizeof(*parent));
    double *state = calloc((N + 1) * K, sizeof(*state));
    double *total = calloc(K, sizeof(*total));
    double *one = calloc(M, sizeof(*one));
    double *tmp = calloc(M, sizeof(*tmp));
    uint64_t j, k, m, w = 0, site = 0, mut = 0, mut_end;
    int32_t e, u, c;
    sweep_t s;

    if (!parent || !state || !total || !one || !tmp) {
        ret = ORC_ERR_NO_MEMORY;
        goto out;
    }
    memset(parent, 0xff, (N + 1) * sizeof(*parent));
    for (j = 0; j < n; j++) {
        u = samples[j];
        for (k = 0; k < K; k++) {
            state[(uint64_t) u * K + k] = weights[j * K + k];
            total[k] += weights[j * K + k];
        }
    }
    memset(result, 0, W * M * sizeof(double));
    sweep_init(&s, t);
    while (sweep_more(&s)) {
        while ((e = sweep_next_out(&s)) != ORC_NULL) {
            c = t->edge_child[e];
            for (u = t->edge_parent[e]; u != ORC_NULL; u = parent[u]) {
                for (k = 0; k < K; k++) {
                    state[(uint64_t) u * K + k] -= state[(uint64_t) c * K + k];
                }
            }
            parent[c] = ORC_NULL;
        }
        while ((e = sweep_next_in(&s)) != ORC_NULL) {
            c = t->edge_child[e];
            parent[c] = t->edge_parent[e];
            for (u = t->edge_parent[e]; u != ORC_NULL; u = parent[u]) {
                for (k = 0; k < K; k++) {
                    state[(uint64_t) u * K + k] += state[(uint64_t) c * K + k];
                }
            }
        }
        sweep_close(&s);
        /* the sites of this tree: left <= position < right (init_trees, trees.c:243-359) */
        while (site < t->num_sites && t->site_position[site] < s.right) {
            while (mut < t->num_mutations && (uint64_t) t->mutation_site[mut] < site) {
                mut++;
            }
            mut_end = mut;
            while (mut_end < t->num_mutations && (uint64_t) t->mutation_site[mut_end] == site) {
                mut_end++;
            }
            ret = site_result(
                t, site, mut, mut_end, state, K, total, M, f, fp, polarised, one, tmp);
            if (ret != 0) {
                goto out;
            }
            while (windows[w + 1] <= t->site_position[site]) {
                w++;
            }
            for (m = 0; m < M; m++) {
                result[w * M + m] += one[m];
            }
            mut = mut_end;
            site++;
        }
        s.left = s.right;
    }
out:
    free(parent);
    free(state);
    free(total);
    free(one);
    free(tmp);
    return ret;
}

/* tsk_treeseq_general_stat, c/tskit/trees.c:2035-2095 (node mode is outside the path) */
int
orc_general_stat(const orc_tables_t *t, uint64_t K, const double *weights, uint64_t M,
    orc_summary_func *f, void *fp, uint64_t W, const double *windows, uint32_t options,
    double *result)
{
    int ret = 0;
    int site = !!(options & ORC_STAT_SITE), branch = !!(options & ORC_STAT_BRANCH),
        node = !!(options & ORC_STAT_NODE);
    double whole[2] = { 0, t->sequence_length };
    uint64_t n = 0, j, k, w, m;
    int32_t *map = NULL, *samples = NULL;
    unpolarised_t up;
    double *total = NULL;

    memset(&up, 0, sizeof(up));
    if (!(site || branch || node)) {
        site = 1;
    }
    if (site + branch + node > 1) {
        return ORC_ERR_MULTIPLE_STAT_MODES;
    }
    if (K < 1) {
        return ORC_ERR_BAD_STATE_DIMS;
    }
    if (M < 1) {
        return ORC_ERR_BAD_RESULT_DIMS;
    }
    if (windows == NULL) {
        W = 1;
        windows = whole;
    } else {
        ret = check_windows(t, W, windows, 1);
        if (ret != 0) {
            return ret;
        }
    }
    if (node) {
        return ORC_ERR_UNSUPPORTED_STAT_MODE;
    }
    map = sample_index_map(t, &n);
    samples = malloc((n + 1) * sizeof(*samples));
    total = calloc(K, sizeof(*total));
    up.tmp_state = calloc(K, sizeof(double));
    up.tmp_result = calloc(M, sizeof(double));
    if (!map || !samples || !total || !up.tmp_state || !up.tmp_result) {
        ret = ORC_ERR_NO_MEMORY;
        goto out;
    }
    for (j = 0; j < t->num_nodes; j++) {
        if (map[j] != ORC_NULL) {
            samples[map[j]] = (int32_t) j;
        }
    }
    if (site) {
        ret = site_stat(t, n, samples, K, weights, M, f, fp, W, windows,
            !!(options & ORC_STAT_POLARISED), result);
    } else {
        if (t->time_uncalibrated && !(options & ORC_STAT_ALLOW_TIME_UNCALIBRATED)) {
            ret = ORC_ERR_TIME_UNCALIBRATED;
            goto out;
        }
        if (options & ORC_STAT_POLARISED) {
            ret = branch_stat(t, n, samples, K, weights, M, f, fp, W, windows, result);
        } else {
            for (j = 0; j < n; j++) {
                for (k = 0; k < K; k++) {
                    total[k] += weights[j * K + k];
                }
            }
            up.f = f;
            up.f_params = fp;
            up.K = K;
            up.total = total;
            ret = branch_stat(t, n, samples, K, weights, M, unpolarised, &up, W, windows, result);
        }
    }
    if (ret == 0 && (options & ORC_STAT_SPAN_NORMALISE)) {
        /* c/tskit/trees.c:1920-1934 */
        for (w = 0; w < W; w++) {
            for (m = 0; m < M; m++) {
                result[w * M + m] /= windows[w + 1] - windows[w];
            }
        }
    }
out:
    free(map);
    free(samples);
    free(total);
    free(up.tmp_state);
    free(up.tmp_result);
    return ret;
}

/* ------------------------------------------------- sample-count statistics */

typedef struct {
    uint64_t K;
    const uint64_t *n; /* sample set sizes */
    const int32_t *idx; /* index tuples */
} count_params_t;

#define SIZE(p, i) ((double) (p)->n[i])

/* c/tskit/trees.c:3934-3948 */
static int
f_diversity(uint64_t K, const double *x, uint64_t M, double *r, void *params)
{
    const count_params_t *p = params;
    uint64_t j;
    (void) M;
    for (j = 0; j < K; j++) {
        double n = SIZE(p, j);
        r[j] = x[j] * (n - x[j]) / (n * (n - 1));
    }
    return 0;
}

/* c/tskit/trees.c:4221-4236 */
static int
f_segregating_sites(uint64_t K, const double *x, uint64_t M, double *r, void *params)
{
    const count_params_t *p = params;
    uint64_t j;
    (void) M;
    for (j = 0; j < K; j++) {
        double n = SIZE(p, j);
        r[j] = (x[j] > 0) * (1 - x[j] / n);
    }
    return 0;
}

/* c/tskit/trees.c:4248-4264 */
static int
f_Y1(uint64_t K, const double *x, uint64_t M, double *r, void *params)
{
    const count_params_t *p = params;
    uint64_t i;
    (void) K;
    for (i = 0; i < M; i++) {
        double ni = SIZE(p, i);
        double denom = ni * (ni - 1) * (ni - 2);
        double numer = x[i] * (ni - x[i]) * (ni - x[i] - 1);
        r[i] = numer / denom;
    }
    return 0;
}

/* c/tskit/trees.c:4690-4709 */
static int
f_divergence(uint64_t K, const double *x, uint64_t M, double *r, void *params)
{
    const count_params_t *p = params;
    uint64_t q;
    (void) K;
    for (q = 0; q < M; q++) {
        int32_t i = p->idx[2 * q], j = p->idx[2 * q + 1];
        double ni = SIZE(p, i), nj = SIZE(p, j);
        double denom = ni * (nj - (i == j));
        r[q] = x[i] * (nj - x[j]) / denom;
    }
    return 0;
}

/* c/tskit/trees.c:4729-4753 */
static int
f_relatedness(uint64_t K, const double *x, uint64_t M, double *r, void *params)
{
    const count_params_t *p = params;
    uint64_t q, k;
    double sumx = 0, meanx;
    for (k = 0; k < K; k++) {
        sumx += x[k] / SIZE(p, k);
    }
    meanx = sumx / (double) K;
    for (q = 0; q < M; q++) {
        int32_t i = p->idx[2 * q], j = p->idx[2 * q + 1];
        r[q] = (x[i] / SIZE(p, i) - meanx) * (x[j] / SIZE(p, j) - meanx);
    }
    return 0;
}

/* c/tskit/trees.c:4755-4773 */
static int
f_relatedness_noncentred(uint64_t K, const double *x, uint64_t M, double *r, void *params)
{
    const count_params_t *p = params;
    uint64_t q;
    (void) K;
    for (q = 0; q < M; q++) {
        int32_t i = p->idx[2 * q], j = p->idx[2 * q + 1];
        r[q] = x[i] * x[j] / (SIZE(p, i) * SIZE(p, j));
    }
    return 0;
}

/* c/tskit/trees.c:4899-4918 */
static int
f_Y2(uint64_t K, const double *x, uint64_t M, double *r, void *params)
{
    const count_params_t *p = params;
    uint64_t q;
    (void) K;
    for (q = 0; q < M; q++) {
        int32_t i = p->idx[2 * q], j = p->idx[2 * q + 1];
        double ni = SIZE(p, i), nj = SIZE(p, j);
        double denom = ni * nj * (nj - 1);
        r[q] = x[i] * (nj - x[j]) * (nj - x[j] - 1) / denom;
    }
    return 0;
}

/* c/tskit/trees.c:4938-4959 */
static int
f_f2(uint64_t K, const double *x, uint64_t M, double *r, void *params)
{
    const count_params_t *p = params;
    uint64_t q;
    (void) K;
    for (q = 0; q < M; q++) {
        int32_t i = p->idx[2 * q], j = p->idx[2 * q + 1];
        double ni = SIZE(p, i), nj = SIZE(p, j);
        double denom = ni * (ni - 1) * nj * (nj - 1);
        double numer = x[i] * (x[i] - 1) * (nj - x[j]) * (nj - x[j] - 1)
                       - x[i] * (ni - x[i]) * (nj - x[j]) * x[j];
        r[q] = numer / denom;
    }
    return 0;
}

/* c/tskit/trees.c:5177-5199 */
static int
f_Y3(uint64_t K, const double *x, uint64_t M, double *r, void *params)
{
    const count_params_t *p = params;
    uint64_t q;
    (void) K;
    for (q = 0; q < M; q++) {
        int32_t i = p->idx[3 * q], j = p->idx[3 * q + 1], k = p->idx[3 * q + 2];
        double ni = SIZE(p, i), nj = SIZE(p, j), nk = SIZE(p, k);
        double denom = ni * nj * nk;
        double numer = x[i] * (nj - x[j]) * (nk - x[k]);
        r[q] = numer / denom;
    }
    return 0;
}

/* c/tskit/trees.c:5219-5242 */
static int
f_f3(uint64_t K, const double *x, uint64_t M, double *r, void *params)
{
    const count_params_t *p = params;
    uint64_t q;
    (void) K;
    for (q = 0; q < M; q++) {
        int32_t i = p->idx[3 * q], j = p->idx[3 * q + 1], k = p->idx[3 * q + 2];
        double ni = SIZE(p, i), nj = SIZE(p, j), nk = SIZE(p, k);
        double denom = ni * (ni - 1) * nj * nk;
        double numer = x[i] * (x[i] - 1) * (nj - x[j]) * (nk - x[k])
                       - x[i] * (ni - x[i]) * (nj - x[j]) * x[k];
        r[q] = numer / denom;
    }
    return 0;
}

/* c/tskit/trees.c:5266-5291 */
static int
f_f4(uint64_t K, const double *x, uint64_t M, double *r, void *params)
{
    const count_params_t *p = params;
    uint64_t q;
    (void) K;
    for (q = 0; q < M; q++) {
        int32_t i = p->idx[4 * q], j = p->idx[4 * q + 1], k = p->idx[4 * q + 2],
                l = p->idx[4 * q + 3];
        double ni = SIZE(p, i), nj = SIZE(p, j), nk = SIZE(p, k), nl = SIZE(p, l);
        double denom = ni * nj * nk * nl;
        double numer = x[i] * x[k] * (nj - x[j]) * (nl - x[l])
                       - x[i] * x[l] * (nj - x[j]) * (nk - x[k]);
        r[q] = numer / denom;
    }
    return 0;
}

/* stat ids shared with the test-suite:
 * 0 diversity 1 segregating_sites 2 Y1 3 divergence 4 Y2 5 f2 6 genetic_relatedness
 * 7 Y3 8 f3 9 f4 */
static const struct {
    orc_summary_func *f;
    int tuple;
} STATS[] = { { f_diversity, 0 }, { f_segregating_sites, 0 }, { f_Y1, 0 }, { f_divergence, 2 },
    { f_Y2, 2 }, { f_f2, 2 }, { f_relatedness, 2 }, { f_Y3, 3 }, { f_f3, 3 }, { f_f4, 4 } };

/* tsk_treeseq_sample_count_stat, c/tskit/trees.c:2174-2220, with the k-way
 * input checks of c/tskit/trees.c:4667-4688 in front */
int
orc_sample_count_stat(const orc_tables_t *t, int stat_id, uint64_t K, const uint64_t *sizes,
    const int32_t *sets, uint64_t num_tuples, const int32_t *tuples, uint64_t W,
    const double *windows, uint32_t options, double *result)
{
    int ret = 0;
    uint64_t n = 0, j, k, l, M;
    int32_t *map = NULL;
    double *weights = NULL;
    count_params_t params = { K, sizes, tuples };
    orc_summary_func *f;
    int tuple;

    if (stat_id < 0 || stat_id > 9) {
        return -1;
    }
    f = STATS[stat_id].f;
    tuple = STATS[stat_id].tuple;
    if (stat_id == 6 && (options & ORC_STAT_NONCENTRED)) {
        f = f_relatedness_noncentred;
    }
    M = K;
    if (tuple > 0) {
        if (K < 1) {
            return ORC_ERR_INSUFFICIENT_SAMPLE_SETS;
        }
        if (num_tuples < 1) {
            return ORC_ERR_INSUFFICIENT_INDEX_TUPLES;
        }
        for (j = 0; j < num_tuples * (uint64_t) tuple; j++) {
            if (tuples[j] < 0 || tuples[j] >= (int32_t) K) {
                return ORC_ERR_BAD_SAMPLE_SET_INDEX;
            }
        }
        M = num_tuples;
    }
    /* c/tskit/trees.c:2114-2149 */
    if (K == 0) {
        return ORC_ERR_INSUFFICIENT_SAMPLE_SETS;
    }
    map = sample_index_map(t, &n);
    if (map == NULL) {
        return ORC_ERR_NO_MEMORY;
    }
    j = 0;
    for (k = 0; k < K; k++) {
        if (sizes[k] == 0) {
            ret = ORC_ERR_EMPTY_SAMPLE_SET;
            goto out;
        }
        for (l = 0; l < sizes[k]; l++, j++) {
            if (sets[j] < 0 || sets[j] >= (int32_t) t->num_nodes) {
                ret = ORC_ERR_NODE_OUT_OF_BOUNDS;
                goto out;
            }
            if (map[sets[j]] == ORC_NULL) {
                ret = ORC_ERR_BAD_SAMPLES;
                goto out;
            }
        }
    }
    weights = calloc((n + 1) * K, sizeof(*weights));
    if (weights == NULL) {
        ret = ORC_ERR_NO_MEMORY;
        goto out;
    }
    j = 0;
    for (k = 0; k < K; k++) {
        for (l = 0; l < sizes[k]; l++, j++) {
            double *cell = weights + (uint64_t) map[sets[j]] * K + k;
            if (*cell != 0) {
                ret = ORC_ERR_DUPLICATE_SAMPLE;
                goto out;
            }
            *cell = 1;
        }
    }
    ret = orc_general_stat(t, K, weights, M, f, &params, W, windows, options, result);
out:
    free(map);
    free(weights);
    return ret;
}

/* ---------------------------------------------------------- trees at x
 * parent array and per-node tracked-sample counts of the tree covering each
 * position, by replaying the sweep (tsk_tree_next semantics,
 * c/tskit/trees.c:6679-6835).  positions must be sorted ascending. */
int
orc_trees_at(const orc_tables_t *t, uint64_t nq, const double *positions,
    const int32_t *tracked, uint64_t num_tracked, int32_t *out_parent, int32_t *out_count)
{
    const uint64_t N = t->num_nodes;
    int32_t *parent = malloc((N + 1) * sizeof(*parent));
    int32_t *count = calloc(N + 1, sizeof(*count));
    uint64_t q = 0, j;
    int32_t e, u, c;
    sweep_t s;

    if (!parent || !count) {
        free(parent);
        free(count);
        return ORC_ERR_NO_MEMORY;
    }
    memset(parent, 0xff, (N + 1) * sizeof(*parent));
    if (tracked == NULL) {
        for (j = 0; j < N; j++) {
            count[j] = (t->node_flags[j] & 1u) ? 1 : 0;
        }
    } else {
        for (j = 0; j < num_tracked; j++) {
            count[tracked[j]] = 1;
        }
    }
    sweep_init(&s, t);
    while (sweep_more(&s) && q < nq) {
        while ((e = sweep_next_out(&s)) != ORC_NULL) {
            c = t->edge_child[e];
            for (u = t->edge_parent[e]; u != ORC_NULL; u = parent[u]) {
                count[u] -= count[c];
            }
            parent[c] = ORC_NULL;
        }
        while ((e = sweep_next_in(&s)) != ORC_NULL) {
            c = t->edge_child[e];
            parent[c] = t->edge_parent[e];
            for (u = parent[c]; u != ORC_NULL; u = parent[u]) {
                count[u] += count[c];
            }
        }
        sweep_close(&s);
        while (q < nq && positions[q] < s.right) {
            memcpy(out_parent + q * N, parent, N * sizeof(int32_t));
            memcpy(out_count + q * N, count, N * sizeof(int32_t));
            q++;
        }
        s.left = s.right;
    }
    free(parent);
    free(count);
    return 0;
}

/* ---------------------------------------------------------- genotypes
 * tsk_variant_decode for every site, c/tskit/genotypes.c:473-594: genotype =
 * allele index; 0 is the ancestral state, new alleles numbered in order of
 * first appearance; mutations applied in table order, each overwriting the
 * samples below its node; isolated samples are -1 (missing) unless
 * ISOLATED_NOT_MISSING.  `samples` lists the nodes to genotype (any nodes).
 * out is [num_sites x num_samples] int32. */
int
orc_genotype_matrix(const orc_tables_t *t, const int32_t *samples, uint64_t num_samples,
    uint32_t options, int32_t *out)
{
    const uint64_t N = t->num_nodes;
    int32_t *parent = malloc((N + 1) * sizeof(*parent));
    int32_t *nchild = calloc(N + 1, sizeof(*nchild));
    uint64_t site = 0, mut = 0, j, a;
    int32_t e, u;
    sweep_t s;
    int ret = 0;

    if (!parent || !nchild) {
        free(parent);
        free(nchild);
        return ORC_ERR_NO_MEMORY;
    }
    memset(parent, 0xff, (N + 1) * sizeof(*parent));
    sweep_init(&s, t);
    while (sweep_more(&s)) {
        while ((e = sweep_next_out(&s)) != ORC_NULL) {
            parent[t->edge_child[e]] = ORC_NULL;
            nchild[t->edge_parent[e]]--;
        }
        while ((e = sweep_next_in(&s)) != ORC_NULL) {
            parent[t->edge_child[e]] = t->edge_parent[e];
            nchild[t->edge_parent[e]]++;
        }
        sweep_close(&s);
        while (site < t->num_sites && t->site_position[site] < s.right) {
            int32_t *g = out + site * num_samples;
            uint64_t m0, m1, num_alleles = 1;
            const char *alleles[256];
            uint64_t allele_len[256];
            while (mut < t->num_mutations && (uint64_t) t->mutation_site[mut] < site) {
                mut++;
            }
            m0 = mut;
            m1 = m0;
            while (m1 < t->num_mutations && (uint64_t) t->mutation_site[m1] == site) {
                m1++;
            }
            alleles[0] = t->site_ancestral_state + t->site_ancestral_state_offset[site];
            allele_len[0] = t->site_ancestral_state_offset[site + 1]
                            - t->site_ancestral_state_offset[site];
            for (j = 0; j < num_samples; j++) {
                u = samples[j];
                g[j] = 0;
                if (!(options & ORC_ISOLATED_NOT_MISSING) && parent[u] == ORC_NULL
                    && nchild[u] == 0) {
                    g[j] = -1;
                }
            }
            for (j = m0; j < m1; j++) {
                const char *der = t->mutation_derived_state + t->mutation_derived_state_offset[j];
                uint64_t len = t->mutation_derived_state_offset[j + 1]
                               - t->mutation_derived_state_offset[j];
                uint64_t q;
                for (a = 0; a < num_alleles; a++) {
                    if (allele_len[a] == len && memcmp(alleles[a], der, len) == 0) {
                        break;
                    }
                }
                if (a == num_alleles) {
                    if (num_alleles == 256) {
                        ret = -1;
                        goto out;
                    }
                    alleles[a] = der;
                    allele_len[a] = len;
                    num_alleles++;
                }
                /* every listed node at or below the mutation's node takes the allele */
                for (q = 0; q < num_samples; q++) {
                    for (u = samples[q]; u != ORC_NULL; u = parent[u]) {
                        if (u == t->mutation_node[j]) {
                            g[q] = (int32_t) a;
                            break;
                        }
                    }
                }
            }
            mut = m1;
            site++;
        }
        s.left = s.right;
    }
out:
    free(parent);
    free(nchild);
    return ret;
}

/* ------------------------------------------------------ divergence matrix
 * tsk_treeseq_divergence_matrix, c/tskit/trees.c:8901-9001, by definition:
 * branch: D[j,k] += (t_mrca - t_u + t_mrca - t_v) * span over sample pairs
 *         (no MRCA: each node's distance to its own root), :8579-8676
 * site:   per site, pairs of samples carrying different alleles, :8684-8826
 * normalised by n_j * n_k (diagonal n_j (n_j - 1)), :8876-8899.
 * sample_sets == NULL: every sample is its own set. */
int
orc_divergence_matrix(const orc_tables_t *t, uint64_t num_sets, const uint64_t *sizes,
    const int32_t *sets, uint64_t W, const double *windows, uint32_t options, double *result)
{
    int ret = 0;
    const uint64_t N = t->num_nodes;
    int site = !!(options & ORC_STAT_SITE), branch = !!(options & ORC_STAT_BRANCH),
        node = !!(options & ORC_STAT_NODE);
    double whole[2] = { 0, t->sequence_length };
    uint64_t n_all = 0, n = 0, j, k, w, a, b;
    int32_t *map = sample_index_map(t, &n_all);
    int32_t *set_of = NULL, *nodes = NULL, *parent = NULL, *geno = NULL;
    uint64_t *set_size = NULL;
    int32_t e, u, v;
    sweep_t s;

    if (map == NULL) {
        return ORC_ERR_NO_MEMORY;
    }
    if (node) {
        ret = ORC_ERR_UNSUPPORTED_STAT_MODE;
        goto out;
    }
    if (!(site || branch)) {
        site = 1;
    }
    if (site + branch > 1) {
        ret = ORC_ERR_MULTIPLE_STAT_MODES;
        goto out;
    }
    if (options & ORC_STAT_POLARISED) {
        ret = ORC_ERR_STAT_POLARISED_UNSUPPORTED;
        goto out;
    }
    if (windows == NULL) {
        W = 1;
        windows = whole;
    } else {
        ret = check_windows(t, W, windows, 0);
        if (ret != 0) {
            goto out;
        }
    }
    set_size = calloc(num_sets + 1, sizeof(*set_size));
    set_of = malloc((N + 1) * sizeof(*set_of));
    parent = malloc((N + 1) * sizeof(*parent));
    if (!set_size || !set_of || !parent) {
        ret = ORC_ERR_NO_MEMORY;
        goto out;
    }
    memset(set_of, 0xff, (N + 1) * sizeof(*set_of));
    if (sets == NULL) {
        if (num_sets != n_all) {
            ret = -1;
            goto out;
        }
        n = n_all;
        nodes = malloc((n + 1) * sizeof(*nodes));
        for (j = 0; j < N; j++) {
            if (map[j] != ORC_NULL) {
                nodes[map[j]] = (int32_t) j;
                set_of[j] = map[j];
                set_size[map[j]] = 1;
            }
        }
    } else {
        for (k = 0; k < num_sets; k++) {
            n += sizes ? sizes[k] : 1;
        }
        nodes = malloc((n + 1) * sizeof(*nodes));
        j = 0;
        for (k = 0; k < num_sets; k++) {
            uint64_t sz = sizes ? sizes[k] : 1;
            set_size[k] = sz;
            for (a = 0; a < sz; a++, j++) {
                u = sets[j];
                if (u < 0 || u >= (int32_t) N) {
                    ret = ORC_ERR_NODE_OUT_OF_BOUNDS;
                    goto out;
                }
                if (!(t->node_flags[u] & 1u)) {
                    ret = ORC_ERR_BAD_SAMPLES;
                    goto out;
                }
                if (set_of[u] != ORC_NULL) {
                    ret = ORC_ERR_DUPLICATE_SAMPLE;
                    goto out;
                }
                set_of[u] = (int32_t) k;
                nodes[j] = u;
            }
        }
    }
    memset(result, 0, W * num_sets * num_sets * sizeof(double));
    if (branch && t->time_uncalibrated && !(options & ORC_STAT_ALLOW_TIME_UNCALIBRATED)) {
        ret = ORC_ERR_TIME_UNCALIBRATED;
        goto out;
    }
    if (site) {
        geno = malloc((t->num_sites * n + 1) * sizeof(*geno));
        if (geno == NULL) {
            ret = ORC_ERR_NO_MEMORY;
            goto out;
        }
        ret = orc_genotype_matrix(t, nodes, n, ORC_ISOLATED_NOT_MISSING, geno);
        if (ret != 0) {
            goto out;
        }
        w = 0;
        for (j = 0; j < t->num_sites; j++) {
            double x = t->site_position[j];
            const int32_t *g = geno + j * n;
            if (x < windows[0] || x >= windows[W]) {
                continue;
            }
            while (windows[w + 1] <= x) {
                w++;
            }
            for (a = 0; a < n; a++) {
                for (b = a + 1; b < n; b++) {
                    if (g[a] != g[b]) {
                        int32_t sa = set_of[nodes[a]], sb = set_of[nodes[b]];
                        result[(w * num_sets + sa) * num_sets + sb] += 1;
                        result[(w * num_sets + sb) * num_sets + sa] += 1;
                    }
                }
            }
        }
    } else {
        memset(parent, 0xff, (N + 1) * sizeof(*parent));
        sweep_init(&s, t);
        w = 0;
        while (sweep_more(&s)) {
            while ((e = sweep_next_out(&s)) != ORC_NULL) {
                parent[t->edge_child[e]] = ORC_NULL;
            }
            while ((e = sweep_next_in(&s)) != ORC_NULL) {
                parent[t->edge_child[e]] = t->edge_parent[e];
            }
            sweep_close(&s);
            for (w = 0; w < W; w++) {
                double lo = s.left > windows[w] ? s.left : windows[w];
                double hi = s.right < windows[w + 1] ? s.right : windows[w + 1];
                if (hi <= lo) {
                    continue;
                }
                for (a = 0; a < n; a++) {
                    for (b = a + 1; b < n; b++) {
                        int32_t sa = set_of[nodes[a]], sb = set_of[nodes[b]];
                        double ta = t->node_time[nodes[a]], tb = t->node_time[nodes[b]];
                        double d;
                        int32_t mrca = ORC_NULL, ra = nodes[a], rb = nodes[b];
                        /* MRCA by marking a's path with time order */
                        u = nodes[a];
                        v = nodes[b];
                        while (u != v && u != ORC_NULL && v != ORC_NULL) {
                            if (t->node_time[u] < t->node_time[v]
                                || (t->node_time[u] == t->node_time[v] && u < v)) {
                                ra = u;
                                u = parent[u];
                            } else {
                                rb = v;
                                v = parent[v];
                            }
                        }
                        if (u == v && u != ORC_NULL) {
                            mrca = u;
                        }
                        if (mrca != ORC_NULL) {
                            d = (t->node_time[mrca] - ta) + (t->node_time[mrca] - tb);
                        } else {
                            /* distinct roots: distance of each to its own root */
                            while (u != ORC_NULL) {
                                ra = u;
                                u = parent[u];
                            }
                            while (v != ORC_NULL) {
                                rb = v;
                                v = parent[v];
                            }
                            d = (t->node_time[ra] - ta) + (t->node_time[rb] - tb);
                        }
                        result[(w * num_sets + sa) * num_sets + sb] += d * (hi - lo);
                        result[(w * num_sets + sb) * num_sets + sa] += d * (hi - lo);
                    }
                }
            }
            s.left = s.right;
        }
    }
    /* c/tskit/trees.c:8876-8899 and span_normalise */
    for (w = 0; w < W; w++) {
        for (j = 0; j < num_sets; j++) {
            for (k = 0; k < num_sets; k++) {
                double denom = (double) set_size[j] * (double) set_size[k];
                double *cell = &result[(w * num_sets + j) * num_sets + k];
                if (j == k) {
                    denom = (double) set_size[j] * ((double) set_size[j] - 1);
                    /* same-set pairs were added to [j,j] twice (once per order) */
                }
                if (denom != 0) {
                    *cell /= denom;
                }
                if (options & ORC_STAT_SPAN_NORMALISE) {
                    *cell /= windows[w + 1] - windows[w];
                }
            }
        }
    }
out:
    free(map);
    free(set_of);
    free(nodes);
    free(parent);
    free(geno);
    free(set_size);
    return ret;
}
