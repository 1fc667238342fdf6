/*
 * tskit_b200.h -- C ABI of the B200 general-statistics engine.
 *
 * Drop-in boundary for tskit's general-statistics hot path.  Every entry point
 * keeps the argument order, types and error codes of the reference function it
 * replaces (cited per function; paths are relative to the tskit repository),
 * except that the `const tsk_treeseq_t *self` argument becomes an opaque
 * handle owning the device copy of the tables ("plan").  Plain pointers and
 * sizes only; all pointers are HOST pointers borrowed for the duration of the
 * call; `result` is caller-allocated and zeroed/overwritten by the callee
 * exactly as the reference does (c/tskit/trees.c:1417, 1703, 8979).
 *
 * Types mirror c/tskit/core.h:99-123: tsk_id_t=int32_t, tsk_size_t=uint64_t,
 * tsk_flags_t=uint32_t, TSK_NULL=-1.
 */
#ifndef TSKIT_B200_H
#define TSKIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- option bits: identical values to c/tskit/trees.h:47-56 ---- */
#define TSKB_STAT_SITE (1u << 0)
#define TSKB_STAT_BRANCH (1u << 1)
#define TSKB_STAT_NODE (1u << 2)
#define TSKB_STAT_POLARISED (1u << 10)
#define TSKB_STAT_SPAN_NORMALISE (1u << 11)
#define TSKB_STAT_ALLOW_TIME_UNCALIBRATED (1u << 12)
#define TSKB_STAT_PAIR_NORMALISE (1u << 13)
#define TSKB_STAT_NONCENTRED (1u << 14)

/* tskb_treeseq_init option: keep the pieces of nodes without a branch above them too, which
 * mode="node" statistics need (every node has a value); costs ~16 % in the other modes. */
#define TSKB_INIT_NODE_MODE (1u << 0)
/* c/tskit/genotypes.h:35 */
#define TSKB_ISOLATED_NOT_MISSING (1u << 1)

/* ---- error codes: identical values to c/tskit/core.h:259-698 ---- */
#define TSKB_ERR_NO_MEMORY (-2)
#define TSKB_ERR_BAD_PARAM_VALUE (-4)
/* table integrity, checked by tskb_treeseq_init as tsk_treeseq_init does through
 * tsk_table_collection_check_integrity (c/tskit/tables.c:10362-10640, 10894-10930) */
#define TSKB_ERR_BAD_OFFSET (-200)
#define TSKB_ERR_NODE_OUT_OF_BOUNDS (-202)
#define TSKB_ERR_EDGE_OUT_OF_BOUNDS (-203)
#define TSKB_ERR_SITE_OUT_OF_BOUNDS (-205)
#define TSKB_ERR_MUTATION_OUT_OF_BOUNDS (-206)
#define TSKB_ERR_TIME_NONFINITE (-210)
#define TSKB_ERR_GENOME_COORDS_NONFINITE (-211)
#define TSKB_ERR_NULL_PARENT (-300)
#define TSKB_ERR_NULL_CHILD (-301)
#define TSKB_ERR_BAD_NODE_TIME_ORDERING (-306)
#define TSKB_ERR_BAD_EDGE_INTERVAL (-307)
#define TSKB_ERR_RIGHT_GREATER_SEQ_LENGTH (-309)
#define TSKB_ERR_LEFT_LESS_ZERO (-310)
#define TSKB_ERR_UNSORTED_SITES (-400)
#define TSKB_ERR_DUPLICATE_SITE_POSITION (-401)
#define TSKB_ERR_BAD_SITE_POSITION (-402)
#define TSKB_ERR_MUTATION_PARENT_DIFFERENT_SITE (-500)
#define TSKB_ERR_MUTATION_PARENT_EQUAL (-501)
#define TSKB_ERR_MUTATION_PARENT_AFTER_CHILD (-502)
#define TSKB_ERR_UNSORTED_MUTATIONS (-504)
#define TSKB_ERR_BAD_SEQUENCE_LENGTH (-701)
#define TSKB_ERR_DUPLICATE_SAMPLE (-600)
#define TSKB_ERR_BAD_SAMPLES (-601)
#define TSKB_ERR_BAD_NUM_WINDOWS (-900)
#define TSKB_ERR_BAD_WINDOWS (-901)
#define TSKB_ERR_MULTIPLE_STAT_MODES (-902)
#define TSKB_ERR_BAD_STATE_DIMS (-903)
#define TSKB_ERR_BAD_RESULT_DIMS (-904)
#define TSKB_ERR_INSUFFICIENT_SAMPLE_SETS (-905)
#define TSKB_ERR_INSUFFICIENT_INDEX_TUPLES (-906)
#define TSKB_ERR_BAD_SAMPLE_SET_INDEX (-907)
#define TSKB_ERR_EMPTY_SAMPLE_SET (-908)
#define TSKB_ERR_UNSUPPORTED_STAT_MODE (-909)
#define TSKB_ERR_TIME_UNCALIBRATED (-910)
#define TSKB_ERR_STAT_POLARISED_UNSUPPORTED (-911)
#define TSKB_ERR_INSUFFICIENT_WEIGHTS (-913)
#define TSKB_ERR_BAD_TIME_WINDOWS_DIM (-922)
#define TSKB_ERR_BAD_TIME_WINDOWS (-924)
/* engine-specific codes, outside tskit's range */
#define TSKB_ERR_CUDA (-20001)           /* a CUDA runtime call failed */
#define TSKB_ERR_BAD_INDEX_ORDER (-20002) /* edge indexes not in canonical order */
#define TSKB_ERR_UNSUPPORTED (-20003)     /* valid tskit call the engine does not accelerate */
#define TSKB_ERR_NO_DEVICE (-20004)       /* no CUDA device: there is NO CPU fallback */

typedef struct tskb_treeseq tskb_treeseq_t;

/* The columns the path reads; same SoA columns as tsk_table_collection_t
 * (c/tskit/tables.h:309-601) and the derived arrays tsk_treeseq_init builds
 * (c/tskit/trees.c:455-545).  edge_*_order are the reference's edge indexes
 * (c/tskit/tables.c:11392-11459), consumed as given. */
typedef struct {
    double sequence_length;
    int32_t time_uncalibrated; /* tsk_treeseq_t.time_uncalibrated, trees.c:538 */
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
} tskb_tables_t;

/* Replaces tsk_treeseq_init (c/tskit/trees.c:455) for this path: stages the
 * tables in HBM and builds the replay plan.  [range_left, range_right) clips
 * the genome (multi-GPU window sharding); pass 0, sequence_length for all of it.
 * device is the CUDA ordinal. */
int tskb_treeseq_init(tskb_treeseq_t **self, const tskb_tables_t *tables, int device,
    double range_left, double range_right, uint32_t options);
int tskb_treeseq_free(tskb_treeseq_t *self);

/* tsk_strerror (c/tskit/core.c) for the codes above */
const char *tskb_strerror(int err);
/* detail of the last TSKB_ERR_CUDA on this thread */
const char *tskb_last_cuda_error(void);

/* one_way_sample_stat_method, c/tskit/trees.h:1098-1112 */
int tskb_treeseq_diversity(const tskb_treeseq_t *self, uint64_t num_sample_sets,
    const uint64_t *sample_set_sizes, const int32_t *sample_sets, uint64_t num_windows,
    const double *windows, uint32_t options, double *result);
int tskb_treeseq_segregating_sites(const tskb_treeseq_t *self, uint64_t num_sample_sets,
    const uint64_t *sample_set_sizes, const int32_t *sample_sets, uint64_t num_windows,
    const double *windows, uint32_t options, double *result);
int tskb_treeseq_Y1(const tskb_treeseq_t *self, uint64_t num_sample_sets,
    const uint64_t *sample_set_sizes, const int32_t *sample_sets, uint64_t num_windows,
    const double *windows, uint32_t options, double *result);

/* general_sample_stat_method, c/tskit/trees.h:1118-1121, 1189-1239 */
int tskb_treeseq_divergence(const tskb_treeseq_t *self, uint64_t num_sample_sets,
    const uint64_t *sample_set_sizes, const int32_t *sample_sets,
    uint64_t num_index_tuples, const int32_t *index_tuples, uint64_t num_windows,
    const double *windows, uint32_t options, double *result);
int tskb_treeseq_Y2(const tskb_treeseq_t *self, uint64_t num_sample_sets,
    const uint64_t *sample_set_sizes, const int32_t *sample_sets,
    uint64_t num_index_tuples, const int32_t *index_tuples, uint64_t num_windows,
    const double *windows, uint32_t options, double *result);
int tskb_treeseq_f2(const tskb_treeseq_t *self, uint64_t num_sample_sets,
    const uint64_t *sample_set_sizes, const int32_t *sample_sets,
    uint64_t num_index_tuples, const int32_t *index_tuples, uint64_t num_windows,
    const double *windows, uint32_t options, double *result);
int tskb_treeseq_genetic_relatedness(const tskb_treeseq_t *self, uint64_t num_sample_sets,
    const uint64_t *sample_set_sizes, const int32_t *sample_sets,
    uint64_t num_index_tuples, const int32_t *index_tuples, uint64_t num_windows,
    const double *windows, uint32_t options, double *result);
int tskb_treeseq_Y3(const tskb_treeseq_t *self, uint64_t num_sample_sets,
    const uint64_t *sample_set_sizes, const int32_t *sample_sets,
    uint64_t num_index_tuples, const int32_t *index_tuples, uint64_t num_windows,
    const double *windows, uint32_t options, double *result);
int tskb_treeseq_f3(const tskb_treeseq_t *self, uint64_t num_sample_sets,
    const uint64_t *sample_set_sizes, const int32_t *sample_sets,
    uint64_t num_index_tuples, const int32_t *index_tuples, uint64_t num_windows,
    const double *windows, uint32_t options, double *result);
int tskb_treeseq_f4(const tskb_treeseq_t *self, uint64_t num_sample_sets,
    const uint64_t *sample_set_sizes, const int32_t *sample_sets,
    uint64_t num_index_tuples, const int32_t *index_tuples, uint64_t num_windows,
    const double *windows, uint32_t options, double *result);

/* Weighted statistics: fp64 node states = sums of per-sample weights.
 * tsk_treeseq_trait_covariance / _trait_correlation: one_way_weighted_method
 * (c/tskit/trees.h:1056-1061; trees.c:3976-4110), `weights` row-major
 * [num_samples x num_weights], result [num_windows x num_weights].
 * tsk_treeseq_genetic_relatedness_weighted: (trees.h:1079-1082;
 * trees.c:4840-4897), result [num_windows x num_index_tuples]; note the reference's argument
 * order (result before options).  At most 8 state columns (7 weights where a frequency column
 * is appended) per sweep; more columns are computed in batches. */
int tskb_treeseq_trait_covariance(const tskb_treeseq_t *self, uint64_t num_weights,
    const double *weights, uint64_t num_windows, const double *windows, uint32_t options,
    double *result);
int tskb_treeseq_trait_correlation(const tskb_treeseq_t *self, uint64_t num_weights,
    const double *weights, uint64_t num_windows, const double *windows, uint32_t options,
    double *result);
/* tsk_treeseq_trait_linear_model (trees.h:1068-1070; trees.c:4106-4219): `covariates` row-major
 * [num_samples x num_covariates], already orthonormalised as the reference assumes;
 * num_weights + num_covariates + 1 <= 8 state columns, else TSKB_ERR_UNSUPPORTED. */
int tskb_treeseq_trait_linear_model(const tskb_treeseq_t *self, uint64_t num_weights,
    const double *weights, uint64_t num_covariates, const double *covariates, uint64_t num_windows,
    const double *windows, uint32_t options, double *result);
int tskb_treeseq_genetic_relatedness_weighted(const tskb_treeseq_t *self, uint64_t num_weights,
    const double *weights, uint64_t num_index_tuples, const int32_t *index_tuples,
    uint64_t num_windows, const double *windows, double *result, uint32_t options);

/* tsk_treeseq_genetic_relatedness_vector (c/tskit/trees.h:1091-1094; trees.c:10772-10816): the
 * product of the branch-mode genetic relatedness matrix with `weights` (row-major
 * [num_samples x num_weights]) without forming the matrix; result
 * [num_windows x num_focal_nodes x num_weights].  Branch mode only (TSK_STAT_SITE / _NODE:
 * TSKB_ERR_UNSUPPORTED_STAT_MODE, as the reference); `windows` need not span the sequence.
 * Focal nodes that are not samples need an engine built with TSKB_INIT_NODE_MODE
 * (TSKB_ERR_UNSUPPORTED otherwise).  fp64 atomics: equal to the reference within rounding. */
int tskb_treeseq_genetic_relatedness_vector(const tskb_treeseq_t *self, uint64_t num_weights,
    const double *weights, uint64_t num_windows, const double *windows, uint64_t num_focal_nodes,
    const int32_t *focal_nodes, double *result, uint32_t options);

/* tsk_treeseq_allele_frequency_spectrum (c/tskit/trees.h:1112-1116; trees.c:3814-3928), site and
 * branch mode: result [num_windows x prod(sample_set_sizes[k] + 1)], row-major over the sets; folded
 * [num_windows x num_time_windows x prod(...)] with time windows (branch mode; trees.c:3663-3680);
 * folded unless TSK_STAT_POLARISED.  Any number of sample sets up to 64 (beyond 7 the spectrum
 * coordinate travels as one fp64 state column).  Time windows other than {0, inf}, or negative node
 * times, need an engine built with TSKB_INIT_NODE_MODE (TSKB_ERR_UNSUPPORTED otherwise); more than
 * 2e9 output cells: TSKB_ERR_UNSUPPORTED. */
int tskb_treeseq_allele_frequency_spectrum(const tskb_treeseq_t *self, uint64_t num_sample_sets,
    const uint64_t *sample_set_sizes, const int32_t *sample_sets, uint64_t num_windows,
    const double *windows, uint64_t num_time_windows, const double *time_windows, uint32_t options,
    double *result);

/* tsk_treeseq_general_stat (c/tskit/trees.h:1035-1037) for the summary
 * functions a device can evaluate: the callback `f` is replaced by a table.
 * For state_dim == 1, `f_table` is [(max_count + 1) x result_dim] with row c =
 * f(c) for the integer sample count c (weights must be 0/1, as
 * tsk_treeseq_sample_count_stat builds them, trees.c:2174-2220). */
int tskb_treeseq_sample_count_stat_tabulated(const tskb_treeseq_t *self,
    uint64_t num_sample_sets, const uint64_t *sample_set_sizes,
    const int32_t *sample_sets, uint64_t result_dim, uint64_t table_rows,
    const double *f_table, uint64_t num_windows, const double *windows,
    uint32_t options, double *result);

/* tsk_treeseq_general_stat (c/tskit/trees.h:1035-1037; trees.c:2035-2095) with the reference's own
 * callback type general_stat_func_t (trees.h:1032-1033): `weights` row-major [num_samples x
 * state_dim], `f(state_dim, state, result_dim, result, params)` returns 0 or an error code, which
 * aborts the call and is returned (trees.c:1396-1399, 1441).  The engine sweeps first, then calls `f`
 * on the host ONCE PER DISTINCT state vector (branch mode: also at total - state unless
 * TSK_STAT_POLARISED; site mode: the allele states) instead of once per node update, and integrates
 * the tabulated values on the device.  Site and branch mode, state_dim <= 8; node mode and wider
 * states: TSKB_ERR_UNSUPPORTED.  Branch mode needs an engine built with TSKB_INIT_NODE_MODE (every
 * piece kept: a summary that is NaN / inf at the state of a node without a branch above it reaches
 * the running sum as 0 x NaN, trees.c:1339-1350); TSKB_ERR_UNSUPPORTED otherwise. */
typedef int tskb_general_stat_func_t(uint64_t state_dim, const double *state, uint64_t result_dim,
    double *result, void *params);
int tskb_treeseq_general_stat(const tskb_treeseq_t *self, uint64_t state_dim, const double *weights,
    uint64_t result_dim, tskb_general_stat_func_t *f, void *f_params, uint64_t num_windows,
    const double *windows, uint32_t options, double *result);

/* tsk_treeseq_divergence_matrix, c/tskit/trees.h:1241 (trees.c:8901-9001) */
int tskb_treeseq_divergence_matrix(const tskb_treeseq_t *self, uint64_t num_sample_sets,
    const uint64_t *sample_set_sizes, const int32_t *sample_sets, uint64_t num_windows,
    const double *windows, uint32_t options, double *result);

/* tsk_variant_decode over all sites (c/tskit/genotypes.c:473-594): int8
 * genotype matrix [num_sites x num_samples] for the given sample nodes
 * (samples == NULL: all samples, in tsk_treeseq_get_samples order). */
int tskb_treeseq_genotype_matrix(const tskb_treeseq_t *self, const int32_t *samples,
    uint64_t num_samples, uint32_t options, int8_t *genotypes);
/* The same decode for the sites [first_site, first_site + num_sites): what a loop of
 * tsk_variant_decode(&var, site_id, 0) over that interval fills (genotypes.c:473-594; the Variant
 * iterator of python/tskit/trees.py:5444-5560 decodes site after site).  int8 [num_sites x
 * num_samples]; TSK_ERR_SITE_OUT_OF_BOUNDS (-205) for an interval outside the site table. */
int tskb_treeseq_decode_sites(const tskb_treeseq_t *self, uint64_t first_site, uint64_t num_sites,
    const int32_t *samples, uint64_t num_samples, uint32_t options, int8_t *genotypes);

/* Integer parity outputs: parent array and per-node tracked-sample counts of
 * the tree covering each position (tsk_tree_seek + tsk_tree_t.parent /
 * tsk_tree_get_num_tracked_samples, trees.c:7092, 6679-6760).  out_* are
 * [num_positions x num_nodes] int32; tracked == NULL counts all samples. */
int tskb_treeseq_trees_at(const tskb_treeseq_t *self, uint64_t num_positions,
    const double *positions, const int32_t *tracked, uint64_t num_tracked,
    int32_t *out_parent, int32_t *out_count);

/* ---- measurement / introspection (no reference counterpart) ---- */
typedef struct {
    uint64_t num_events;       /* edge diffs replayed per sweep */
    uint64_t num_visits;       /* sum over edge diffs of ancestors visited (d-bar * events) */
    uint64_t num_levels;       /* dependent steps of the sweep (piece heights) */
    double stage_ms;           /* wall time of tskb_treeseq_init */
    double last_call_ms;       /* device time of the last statistic call (CUDA events) */
    double last_kernel_ms[8];  /* per phase: 0 weights, 1 sweep, 2 summary, 3 scan + window integration (site: window sums), 4 idle, 5 d2h;
                                * after divergence_matrix(site): 0 decode, 1 contraction, 2 normalise + d2h, 7 alleles */
    uint64_t last_launches;    /* kernels launched by the last statistic call */
    uint64_t device_bytes;     /* HBM held by the plan */
} tskb_stats_t;
int tskb_treeseq_get_stats(const tskb_treeseq_t *self, tskb_stats_t *out);

/* Same statistic entry points with inputs/outputs already resident in HBM
 * (device pointers), used by bench.py to time the kernels without PCIe. */
int tskb_treeseq_stat_device(const tskb_treeseq_t *self, int stat_id,
    uint64_t num_sample_sets, const uint64_t *sample_set_sizes,
    const int32_t *d_sample_sets, uint64_t num_index_tuples, const int32_t *index_tuples,
    uint64_t num_windows, const double *windows, uint32_t options, double *d_result);

/* Sum of the ranks' partial results over NVLink peer memory (multi-GPU genome sharding; no reference
 * counterpart: the reference has no multi-device path; the semantics kept are trees.c:1920-1934,
 * normalise after accumulation).  One exchange object per rank (one process per GPU of one node, or
 * several in one process):
 *   create       allocates this rank's receive slab for results of at most `capacity` doubles;
 *   get_handle   the slab's CUDA IPC handle (TSKB_EXCHANGE_HANDLE_BYTES bytes) for the other processes;
 *   connect      maps the peers' slabs: `handles` = world x TSKB_EXCHANGE_HANDLE_BYTES bytes in rank
 *                order (the own entry is ignored); every rank connects before any rank sums;
 *   connect_local  same for exchange objects living in this process (members[world], in rank order);
 *   sum          d_out[i] = sum over ranks r, in rank order, of rank r's d_local[i], divided by
 *                d_spans[(i / span_stride) % span_count] when d_spans != NULL.  All ranks call it the
 *                same number of times with the same count.  DEVICE pointers; d_out may equal d_local.
 *                With `engine` the kernels run on that engine's stream (after the statistic that
 *                produced d_local), else on the exchange's own.  Host-synchronous unless
 *                TSKB_EXCHANGE_ASYNC; TSKB_ERR_CUDA when a peer's partial does not arrive in 20 s;
 *   status       waits for the calls issued so far and reports a timed-out one. */
#define TSKB_EXCHANGE_HANDLE_BYTES 64
#define TSKB_EXCHANGE_ASYNC 1u
typedef struct tskb_exchange tskb_exchange_t;
int tskb_exchange_create(int device, uint64_t capacity, uint32_t world, uint32_t rank, tskb_exchange_t **out);
int tskb_exchange_get_handle(const tskb_exchange_t *self, void *handle_out);
int tskb_exchange_connect(tskb_exchange_t *self, const void *handles);
int tskb_exchange_connect_local(tskb_exchange_t *self, tskb_exchange_t *const *members);
int tskb_exchange_sum(tskb_exchange_t *self, const tskb_treeseq_t *engine, const double *d_local, uint64_t count,
    const double *d_spans, uint64_t span_stride, uint64_t span_count, double *d_out, uint32_t options);
int tskb_exchange_status(tskb_exchange_t *self, const tskb_treeseq_t *engine);
int tskb_exchange_free(tskb_exchange_t *self);

/* Debug/test access to plan arrays (copied to host). name in: "ev_pos",
 * "ev_child", "ev_sign", "voff", "bp_pos", "q_off", "refs", "q_bp0", "q_bp1", "q_bl",
 * "tile_dep", "level", "rank_node", "level_begin", "mut_src", "mut_allele", "mut_alt", "trace".  Returns the element count, or <0 on error; copies at most
 * max_bytes. */
int64_t tskb_treeseq_debug_array(const tskb_treeseq_t *self, const char *name, void *out,
    uint64_t max_bytes);

#ifdef __cplusplus
}
#endif
#endif
