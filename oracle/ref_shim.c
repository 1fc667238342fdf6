/*
 * ref_shim.c -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * A flat-array C entry layer in front of the UNMODIFIED reference C library.
 * It is compiled together with the reference's own sources *where they lie*
 * under /root/reference/c (never copied into this repo) by oracle/build_ref.sh
 * into oracle/_ref/libtskit_ref.so.  Nothing under tskit_b200/ may link, load
 * or call this file; it is used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * Every function here only marshals plain pointers into the reference's public
 * API (c/tskit/tables.h, c/tskit/trees.h, c/tskit/genotypes.h) and returns the
 * reference's own error code.
 */
#include <stdlib.h>
#include <string.h>
#include <tskit.h>

typedef struct {
    tsk_table_collection_t tables;
    tsk_treeseq_t ts;
    int have_tables;
    int have_ts;
} ref_ts_t;

void
ref_treeseq_free(ref_ts_t *self)
{
    if (self == NULL) {
        return;
    }
    if (self->have_ts) {
        tsk_treeseq_free(&self->ts);
    }
    if (self->have_tables) {
        tsk_table_collection_free(&self->tables);
    }
    free(self);
}

/* Build a tsk_treeseq_t from column arrays. Indexes are built by the reference
 * (tsk_table_collection_build_index, c/tskit/tables.c:11392) and mutation
 * parents are computed by the reference when mut_parent == NULL. */
int
ref_treeseq_new(ref_ts_t **out, double sequence_length, int time_uncalibrated,
    tsk_size_t num_nodes, const tsk_flags_t *node_flags, const double *node_time,
    tsk_size_t num_edges, const double *edge_left, const double *edge_right,
    const tsk_id_t *edge_parent, const tsk_id_t *edge_child, tsk_size_t num_sites,
    const double *site_position, const char *ancestral_state,
    const tsk_size_t *ancestral_state_offset, tsk_size_t num_mutations,
    const tsk_id_t *mut_site, const tsk_id_t *mut_node, const tsk_id_t *mut_parent,
    const char *derived_state, const tsk_size_t *derived_state_offset)
{
    int ret;
    tsk_size_t j;
    tsk_id_t *pop = NULL, *ind = NULL, *mpar = NULL;
    double *mtime = NULL;
    tsk_size_t *meta_off = NULL;
    tsk_size_t max_rows = num_nodes;
    tsk_flags_t init_flags = TSK_TS_INIT_BUILD_INDEXES;
    ref_ts_t *self = calloc(1, sizeof(*self));

    *out = NULL;
    if (self == NULL) {
        return TSK_ERR_NO_MEMORY;
    }
    if (num_edges > max_rows) {
        max_rows = num_edges;
    }
    if (num_sites > max_rows) {
        max_rows = num_sites;
    }
    if (num_mutations > max_rows) {
        max_rows = num_mutations;
    }
    ret = tsk_table_collection_init(&self->tables, 0);
    if (ret != 0) {
        goto out;
    }
    self->have_tables = 1;
    self->tables.sequence_length = sequence_length;
    if (time_uncalibrated) {
        ret = tsk_table_collection_set_time_units(&self->tables, "uncalibrated", 12);
    } else {
        ret = tsk_table_collection_set_time_units(&self->tables, "generations", 11);
    }
    if (ret != 0) {
        goto out;
    }
    pop = malloc((num_nodes + 1) * sizeof(*pop));
    ind = malloc((num_nodes + 1) * sizeof(*ind));
    meta_off = calloc(max_rows + 2, sizeof(*meta_off));
    mpar = malloc((num_mutations + 1) * sizeof(*mpar));
    mtime = malloc((num_mutations + 1) * sizeof(*mtime));
    if (pop == NULL || ind == NULL || meta_off == NULL || mpar == NULL
        || mtime == NULL) {
        ret = TSK_ERR_NO_MEMORY;
        goto out;
    }
    for (j = 0; j < num_nodes; j++) {
        pop[j] = TSK_NULL;
        ind[j] = TSK_NULL;
    }
    ret = tsk_node_table_set_columns(&self->tables.nodes, num_nodes, node_flags,
        node_time, pop, ind, NULL, NULL);
    if (ret != 0) {
        goto out;
    }
    ret = tsk_edge_table_set_columns(&self->tables.edges, num_edges, edge_left,
        edge_right, edge_parent, edge_child, NULL, NULL);
    if (ret != 0) {
        goto out;
    }
    if (num_sites > 0) {
        ret = tsk_site_table_set_columns(&self->tables.sites, num_sites, site_position,
            ancestral_state, ancestral_state_offset, NULL, NULL);
        if (ret != 0) {
            goto out;
        }
    }
    if (num_mutations > 0) {
        for (j = 0; j < num_mutations; j++) {
            mpar[j] = mut_parent == NULL ? TSK_NULL : mut_parent[j];
            mtime[j] = TSK_UNKNOWN_TIME;
        }
        ret = tsk_mutation_table_set_columns(&self->tables.mutations, num_mutations,
            mut_site, mut_node, mpar, mtime, derived_state, derived_state_offset, NULL,
            NULL);
        if (ret != 0) {
            goto out;
        }
        if (mut_parent == NULL) {
            init_flags |= TSK_TS_INIT_COMPUTE_MUTATION_PARENTS;
        }
    }
    ret = tsk_treeseq_init(&self->ts, &self->tables, init_flags);
    if (ret != 0) {
        goto out;
    }
    self->have_ts = 1;
    *out = self;
    self = NULL;
out:
    free(pop);
    free(ind);
    free(meta_off);
    free(mpar);
    free(mtime);
    if (self != NULL) {
        ref_treeseq_free(self);
    }
    return ret;
}

const char *
ref_strerror(int err)
{
    return tsk_strerror(err);
}

tsk_size_t
ref_num_trees(const ref_ts_t *self)
{
    return tsk_treeseq_get_num_trees(&self->ts);
}

tsk_size_t
ref_num_samples(const ref_ts_t *self)
{
    return tsk_treeseq_get_num_samples(&self->ts);
}

void
ref_get_samples(const ref_ts_t *self, tsk_id_t *out)
{
    memcpy(out, tsk_treeseq_get_samples(&self->ts),
        tsk_treeseq_get_num_samples(&self->ts) * sizeof(*out));
}

void
ref_get_breakpoints(const ref_ts_t *self, double *out)
{
    memcpy(out, tsk_treeseq_get_breakpoints(&self->ts),
        (tsk_treeseq_get_num_trees(&self->ts) + 1) * sizeof(*out));
}

/* The reference's own edge indexes and mutation parents, so that the product
 * can be handed exactly what `tsk_treeseq_t` holds. */
void
ref_get_indexes(const ref_ts_t *self, tsk_id_t *insertion, tsk_id_t *removal)
{
    /* tsk_treeseq_init copied the tables; the indexes live on its copy */
    tsk_size_t n = self->ts.tables->edges.num_rows;
    memcpy(insertion, self->ts.tables->indexes.edge_insertion_order, n * sizeof(tsk_id_t));
    memcpy(removal, self->ts.tables->indexes.edge_removal_order, n * sizeof(tsk_id_t));
}

void
ref_get_mutation_parents(const ref_ts_t *self, tsk_id_t *out)
{
    memcpy(out, self->ts.tables->mutations.parent,
        self->ts.tables->mutations.num_rows * sizeof(tsk_id_t));
}

/* which: 0 diversity, 1 segregating_sites, 2 Y1 (trees.h:1103-1112) */
int
ref_one_way_stat(const ref_ts_t *self, int which, tsk_size_t num_sample_sets,
    const tsk_size_t *sample_set_sizes, const tsk_id_t *sample_sets,
    tsk_size_t num_windows, const double *windows, tsk_flags_t options, double *result)
{
    switch (which) {
        case 0:
            return tsk_treeseq_diversity(&self->ts, num_sample_sets, sample_set_sizes,
                sample_sets, num_windows, windows, options, result);
        case 1:
            return tsk_treeseq_segregating_sites(&self->ts, num_sample_sets,
                sample_set_sizes, sample_sets, num_windows, windows, options, result);
        case 2:
            return tsk_treeseq_Y1(&self->ts, num_sample_sets, sample_set_sizes,
                sample_sets, num_windows, windows, options, result);
    }
    return -1;
}

/* which: 0 divergence, 1 Y2, 2 f2, 3 genetic_relatedness, 4 Y3, 5 f3, 6 f4
 * (trees.h:1189-1239) */
int
ref_k_way_stat(const ref_ts_t *self, int which, tsk_size_t num_sample_sets,
    const tsk_size_t *sample_set_sizes, const tsk_id_t *sample_sets,
    tsk_size_t num_index_tuples, const tsk_id_t *index_tuples, tsk_size_t num_windows,
    const double *windows, tsk_flags_t options, double *result)
{
    general_sample_stat_method *m = NULL;
    switch (which) {
        case 0:
            m = tsk_treeseq_divergence;
            break;
        case 1:
            m = tsk_treeseq_Y2;
            break;
        case 2:
            m = tsk_treeseq_f2;
            break;
        case 3:
            m = tsk_treeseq_genetic_relatedness;
            break;
        case 4:
            m = tsk_treeseq_Y3;
            break;
        case 5:
            m = tsk_treeseq_f3;
            break;
        case 6:
            m = tsk_treeseq_f4;
            break;
        default:
            return -1;
    }
    return m(&self->ts, num_sample_sets, sample_set_sizes, sample_sets, num_index_tuples,
        index_tuples, num_windows, windows, options, result);
}

int
ref_general_stat(const ref_ts_t *self, tsk_size_t state_dim, const double *weights,
    tsk_size_t result_dim, general_stat_func_t *f, void *f_params,
    tsk_size_t num_windows, const double *windows, tsk_flags_t options, double *result)
{
    return tsk_treeseq_general_stat(&self->ts, state_dim, weights, result_dim, f,
        f_params, num_windows, windows, options, result);
}

int
ref_divergence_matrix(const ref_ts_t *self, tsk_size_t num_sample_sets,
    const tsk_size_t *sample_set_sizes, const tsk_id_t *sample_sets,
    tsk_size_t num_windows, const double *windows, tsk_flags_t options, double *result)
{
    return tsk_treeseq_divergence_matrix(&self->ts, num_sample_sets, sample_set_sizes,
        sample_sets, num_windows, windows, options, result);
}

/* Decode every site for all samples (samples == NULL) or the given node list
 * with tsk_variant_decode (genotypes.c:473). out is [num_sites x n] int32. */
int
ref_genotype_matrix(const ref_ts_t *self, const tsk_id_t *samples,
    tsk_size_t num_samples, tsk_flags_t options, int32_t *out)
{
    int ret;
    tsk_variant_t var;
    tsk_size_t n = samples == NULL ? tsk_treeseq_get_num_samples(&self->ts) : num_samples;
    tsk_size_t num_sites = tsk_treeseq_get_num_sites(&self->ts);
    tsk_size_t j;

    ret = tsk_variant_init(&var, &self->ts, samples, num_samples, NULL, options);
    if (ret != 0) {
        return ret;
    }
    for (j = 0; j < num_sites; j++) {
        ret = tsk_variant_decode(&var, (tsk_id_t) j, 0);
        if (ret != 0) {
            break;
        }
        memcpy(out + j * n, var.genotypes, n * sizeof(int32_t));
    }
    tsk_variant_free(&var);
    return ret;
}

/* Parent array and per-node sample counts of the tree covering each query
 * position (tsk_tree_seek, trees.c:7092; tsk_tree_get_num_samples). With
 * tracked != NULL, counts are of the tracked samples instead. */
int
ref_trees_at(const ref_ts_t *self, tsk_size_t num_positions, const double *positions,
    const tsk_id_t *tracked, tsk_size_t num_tracked, tsk_id_t *out_parent,
    tsk_id_t *out_count)
{
    int ret;
    tsk_tree_t tree;
    tsk_size_t N = self->ts.tables->nodes.num_rows;
    tsk_size_t q, u, c;

    ret = tsk_tree_init(&tree, &self->ts, 0);
    if (ret != 0) {
        return ret;
    }
    if (tracked != NULL) {
        ret = tsk_tree_set_tracked_samples(&tree, num_tracked, tracked);
        if (ret != 0) {
            goto out;
        }
    }
    for (q = 0; q < num_positions; q++) {
        ret = tsk_tree_seek(&tree, positions[q], 0);
        if (ret != 0) {
            goto out;
        }
        for (u = 0; u < N; u++) {
            out_parent[q * N + u] = tree.parent[u];
            if (tracked != NULL) {
                ret = tsk_tree_get_num_tracked_samples(&tree, (tsk_id_t) u, &c);
            } else {
                ret = tsk_tree_get_num_samples(&tree, (tsk_id_t) u, &c);
            }
            if (ret != 0) {
                goto out;
            }
            out_count[q * N + u] = (tsk_id_t) c;
        }
    }
out:
    tsk_tree_free(&tree);
    return ret;
}

/* Sort + simplify a raw (unsorted) node/edge table pair in place, used only to
 * cross-check the in-repo generator against TableCollection.simplify
 * (tables.c:8961+). Returns the new sizes; columns are written back into the
 * caller's arrays (which must have room for the input sizes). */
int
ref_sort_simplify(double sequence_length, tsk_size_t *num_nodes, tsk_flags_t *node_flags,
    double *node_time, tsk_size_t *num_edges, double *edge_left, double *edge_right,
    tsk_id_t *edge_parent, tsk_id_t *edge_child, const tsk_id_t *samples,
    tsk_size_t num_samples, tsk_id_t *node_map)
{
    int ret;
    tsk_size_t j;
    tsk_table_collection_t tables;
    tsk_id_t *pop = malloc((*num_nodes + 1) * sizeof(*pop));
    tsk_size_t max_rows = *num_nodes > *num_edges ? *num_nodes : *num_edges;
    tsk_size_t *meta_off = calloc(max_rows + 2, sizeof(*meta_off));

    ret = tsk_table_collection_init(&tables, 0);
    if (ret != 0) {
        goto out2;
    }
    tables.sequence_length = sequence_length;
    for (j = 0; j < *num_nodes; j++) {
        pop[j] = TSK_NULL;
    }
    ret = tsk_node_table_set_columns(
        &tables.nodes, *num_nodes, node_flags, node_time, pop, pop, NULL, NULL);
    if (ret != 0) {
        goto out;
    }
    ret = tsk_edge_table_set_columns(&tables.edges, *num_edges, edge_left, edge_right,
        edge_parent, edge_child, NULL, NULL);
    if (ret != 0) {
        goto out;
    }
    ret = tsk_table_collection_sort(&tables, NULL, 0);
    if (ret != 0) {
        goto out;
    }
    ret = tsk_table_collection_simplify(&tables, samples, num_samples, 0, node_map);
    if (ret != 0) {
        goto out;
    }
    *num_nodes = tables.nodes.num_rows;
    *num_edges = tables.edges.num_rows;
    memcpy(node_flags, tables.nodes.flags, *num_nodes * sizeof(*node_flags));
    memcpy(node_time, tables.nodes.time, *num_nodes * sizeof(*node_time));
    memcpy(edge_left, tables.edges.left, *num_edges * sizeof(double));
    memcpy(edge_right, tables.edges.right, *num_edges * sizeof(double));
    memcpy(edge_parent, tables.edges.parent, *num_edges * sizeof(tsk_id_t));
    memcpy(edge_child, tables.edges.child, *num_edges * sizeof(tsk_id_t));
out:
    tsk_table_collection_free(&tables);
out2:
    free(pop);
    free(meta_off);
    return ret;
}
