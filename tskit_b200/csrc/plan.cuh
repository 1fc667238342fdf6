// plan.cuh -- the device-resident "replay plan" of one tree sequence.
//
// The reference sweeps the genome left to right applying edge diffs and, for
// every diff, walks from the edge's parent to the root updating per-node state
// (c/tskit/trees.c:1424-1507, 1712-1734).  That sweep is sequential.  The plan
// turns it into data-parallel work once per tree sequence:
//
//   * event i      = one edge diff in the reference's exact processing order
//                    (all removals at x in removal-index order, then all
//                    insertions at x in insertion-index order);
//   * visit (i, u) = one iteration of the reference's `while (u != TSK_NULL)`
//                    walk for event i.  The chain of an event is found without
//                    any sweep state: because removals are ordered old-parent
//                    first and insertions young-parent first
//                    (c/tskit/tables.c:11392-11459), the walk follows exactly
//                    the edges that span ACROSS x (left < x < right), which an
//                    interval query in a child-major edge CSR answers.  The same
//                    ordering means state[child] read by a removal is the child's
//                    state in the tree LEFT of x and by an insertion its state in
//                    the tree RIGHT of x -- never an intermediate one;
//   * piece (u, t) = node u between two consecutive breakpoints that touch it
//                    (u visited by, or the child of, a diff at breakpoint t):
//                    state[u] and branch_length[u] are constant over a piece.
//                    Branch statistics are sums over pieces of
//                    branch_length * f(state) * |piece ^ window|; the reference's
//                    running sum (trees.c:1339-1350) telescopes to exactly this;
//   * references   = state(u, t) is the node's own sample weight plus the states of its
//                    children in the tree right of t: every piece lists the pieces that hold
//                    those (the child's last piece starting at or before t, found at staging
//                    time; the node's own INIT piece for its weight).  A piece is then just a
//                    sum of a few earlier pieces -- no running state, no scan;
//   * heights      = a piece's height is 1 + the largest height among the pieces it references
//                    (its node's height in that tree).  Pieces of one height are independent
//                    once the lower ones are done; there are as many dependent steps as the
//                    tallest tree is high (39 for the C2 workload) instead of one per edge diff.
//                    The "q_" arrays hold the pieces in processing order (height, then node-major
//                    piece id), every height padded to whole tiles; a piece carries the breakpoints
//                    it starts and ends at, so the branch summary adds +G / -G to the
//                    per-breakpoint deltas of the reference's running sum (trees.c:1339-1350)
//                    without looking at any neighbour.
#pragma once

#include <mutex>
#include <vector>

#include "../../include/tskit_b200.h"
#include "common.cuh"

namespace tskb {

#ifndef TSKB_PROP_TILE
#define TSKB_PROP_TILE 1024
#endif
constexpr uint32_t PROP_TILE = TSKB_PROP_TILE;   // pieces per sweep tile
constexpr uint32_t NO_PIECE = 0xffffffffu;  // padding entry of the processing order

struct Plan {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    // sizes
    uint64_t N = 0, E = 0, S = 0, Mu = 0;
    double L = 0, range_left = 0, range_right = 0;
    int time_uncalibrated = 0;
    uint32_t nev = 0;      // events (edge diffs) inside the range
    uint32_t V = 0;        // visits
    uint32_t nlevels = 0;  // max level + 1
    uint32_t T = 0;        // breakpoints (distinct diff positions) inside the range
    uint32_t num_samples = 0;
    uint32_t site_lo = 0, site_hi = 0;  // sites inside the range
    uint64_t total_alleles = 0;
    uint32_t max_muts_per_site = 0, max_alleles_per_site = 1;

    // host copies needed for argument validation
    std::vector<int32_t> sample_index_map;  // node -> sample index or -1 (trees.c:404-453)
    std::vector<int32_t> samples;
    std::vector<uint32_t> level_begin;  // (padded) processing-order offset of each height

    // --- tables in HBM ---
    DevArray<double> time;            // [N]
    DevArray<int32_t> d_samples;      // [n]
    // child-major CSR of all edges, sorted by (child, left)
    DevArray<uint32_t> coff;          // [N + 1]
    DevArray<double> csr_left, csr_right;
    DevArray<int32_t> csr_parent;
    // --- events (edge diffs) ---
    DevArray<double> ev_pos;          // [nev] breakpoint of the diff (clipped to the range)
    DevArray<int32_t> ev_child;       // [nev]
    DevArray<int8_t> ev_sign;         // [nev] -1 removal, +1 insertion
    DevArray<uint32_t> voff;          // [nev + 1] visit offsets
    // --- pieces in processing order: real pieces by (height, node-major piece id), every height
    //     padded to whole tiles (padding: no references, zero branch length).  The state array of
    //     a call is [npp piece states | one INIT slot per sample (its weight) | one zero slot].
    uint32_t P = 0;                   // real + INIT pieces of the node-major construction order
    uint32_t nheights = 0, npp = 0, nrefs = 0, ntiles = 0;
    DevArray<uint32_t> q_bp0;         // [npp] breakpoint (index into bp_pos) the piece starts at
    DevArray<uint32_t> q_bp1;         // [npp] breakpoint it ends at: the node's next piece, or T
                                      //       (= range_right); NO_PIECE marks padding
    DevArray<double> bp_pos;          // [T + 1] distinct diff positions, then range_right
    DevArray<uint32_t> q_eff;         // [npp] breakpoint of the node's last update before the piece (branch AFS;
                                      //       0xffffffff: start of the range)
    bool has_negative_time = false;
    DevArray<double> q_bl;            // [npp] branch length above the node over the piece
    DevArray<uint32_t> q_off;         // [npp + 1] offsets into refs
    DevArray<uint32_t> refs;          // [nrefs] state slots whose values add up to the piece's state
    DevArray<uint32_t> tile_dep;      // [ntiles] first tile of the tile's height = number of tiles
                                      //  that must be complete before its gathers
    DevArray<int32_t> d_sample_index; // [N] node -> sample index or -1
    // --- node mode (TSKB_INIT_NODE_MODE): every piece is kept, and every piece knows its node
    bool all_pieces = false;
    DevArray<int32_t> q_node;         // [npp] node of the piece (-1: padding)
    DevArray<uint32_t> node_first_bp; // [N] breakpoint of the node's first piece (T: it has none)
    // --- parent-major edge CSR, sorted by (parent, left); pmax = running max of right within the
    //     parent's list, which bounds the backward scan of an interval-stabbing query
    DevArray<uint32_t> pm_off;        // [N + 1]
    DevArray<double> pm_left, pm_right, pm_pmax;
    DevArray<int32_t> pm_child;
    DevArray<int32_t> rank_node;      // [N] rank -> node id (nodes sorted by (level, id))
    DevArray<uint32_t> level;         // [N]
    // --- sites ---
    DevArray<double> site_pos;        // [S]
    DevArray<uint32_t> site_moff;     // [S + 1] mutation CSR
    DevArray<uint32_t> site_aoff;     // [S + 1] allele-slot CSR
    DevArray<int32_t> mut_node;       // [Mu]
    DevArray<int32_t> mut_src;        // [Mu] state slot holding state[mutation.node] at the site
    DevArray<uint16_t> mut_allele;    // [Mu] allele index of the derived state
    DevArray<uint16_t> mut_alt;       // [Mu] allele index the mutation's state is subtracted from
    std::vector<double> h_site_pos;

    // --- summary order (built on the first many-column branch call, stats.cu:ensure_summary_order):
    //     the real pieces sorted by the breakpoint they start at, with the state slot of each
    mutable bool so_built = false;
    mutable uint32_t nsp = 0;
    mutable DevArray<uint32_t> so_slot, so_bp0, so_bp1;  // [nsp]
    mutable DevArray<double> so_bl;                      // [nsp]
    // --- positions of every piece's ends (built on the first call of the window-bin summary,
    //     stats.cu:ensure_piece_positions): bp_pos[q_bp0], bp_pos[q_bp1] in processing order
    mutable bool qx_built = false;
    mutable DevArray<double> q_x0, q_x1;                 // [npp]

    // per-call scratch + statistics
    mutable std::mutex mu;
    mutable std::vector<uint32_t> seen_stamp;  // argument validation scratch (api.cu:check_sample_sets)
    mutable uint64_t seen_epoch = 0;
    mutable Arena arena;
    mutable tskb_stats_t stats = {};
    mutable unsigned long long *stats_trace = nullptr;  // TSKB_TRACE: per-tile timeline of the last call (arena)
    size_t scan_temp_bytes = 0;

    ~Plan();
    uint64_t device_bytes() const;
};

// plan.cu
Plan *build_plan(const tskb_tables_t *t, int device, double range_left, double range_right,
    uint32_t options);

// device -> pageable host memory through per-thread pinned staging buffers (large results); returns
// when the copy is complete
void staged_download(int device, void *dst, const void *src, size_t bytes, cudaStream_t stream);

// stats.cu
struct StatSpec {
    int stat_id;          // see StatId
    uint32_t K;           // number of sample sets (state_dim)
    uint32_t M;           // result_dim
    uint32_t tuple;       // index tuple width (0 for one-way)
    const uint64_t *sizes;       // host [K]
    const int32_t *sets;         // flattened sample ids (host or device, see sets_on_device)
    bool sets_on_device;
    const int32_t *indexes;      // host [M * tuple]
    uint32_t W;
    const double *windows;       // host [W + 1]
    uint32_t options;
    double *result;              // host or device [W * M]
    bool result_on_device;
    // weighted statistics (stat_id >= STAT_TRAIT_COV): K columns of per-sample weights, already
    // pre-processed as the reference does (centred / standardised / frequency column appended)
    const double *weights = nullptr;        // host [num_samples * K], row-major
    const double *column_totals = nullptr;  // host [K]
    // tabulated summary (stat_id == STAT_TABULATED)
    const double *f_table = nullptr;
    uint64_t table_rows = 0;
    // STAT_AFS: size of one window's spectrum = product of (set size + 1); result is
    // [W x num_time_windows x afs_size]
    uint64_t afs_size = 0;
    // more than 7 sample sets: fp64 states (linear spectrum coordinate, all-samples count); afs_dims =
    // set sizes + 1 (host [afs_nsets])
    const uint32_t *afs_dims = nullptr;
    uint32_t afs_nsets = 0;
    // branch mode with time windows other than [0, inf) (needs the node of every piece: a node-mode plan)
    const double *time_windows = nullptr;   // host [num_time_windows + 1]; nullptr: the default window
    uint32_t num_time_windows = 1;
    // STAT_REL_VECTOR: focal nodes (host); M = num_focal * K; focal_needs_nodes: some focal node is
    // not a sample (needs the node of every piece: a node-mode plan)
    const int32_t *focal = nullptr;
    uint64_t num_focal = 0;
    bool focal_needs_nodes = false;
};

enum StatId {
    STAT_DIVERSITY = 0, STAT_SEGSITES = 1, STAT_Y1 = 2, STAT_DIVERGENCE = 3, STAT_Y2 = 4,
    STAT_F2 = 5, STAT_RELATEDNESS = 6, STAT_Y3 = 7, STAT_F3 = 8, STAT_F4 = 9,
    STAT_RELATEDNESS_NC = 10, STAT_TABULATED = 11,
    // weighted statistics: fp64 states (trees.c:3960-4110, 4800-4897)
    STAT_TRAIT_COV = 12, STAT_TRAIT_CORR = 13, STAT_REL_WEIGHTED = 14, STAT_REL_WEIGHTED_NC = 15,
    STAT_TRAIT_LM = 16,
    // joint allele frequency spectrum, site mode (trees.c:3497-3648): K sets + the all-samples column
    STAT_AFS = 17,
    // GRM x vector, branch mode (trees.c:10445-10816): fp64 states, result [W x num_focal x K]
    STAT_REL_VECTOR = 18,
    // centred genetic_relatedness over more sample sets than a sweep carries (trees.c:4729-4753): fp64
    // states = indicator columns of the batch's sets + one column summing 1/n_k over ALL sets
    STAT_REL_SIDE = 19
};

// tsk_treeseq_general_stat with a host callback (general_stat_func_t, trees.h:1032-1033)
typedef int (*general_stat_func)(uint64_t state_dim, const double *state, uint64_t result_dim, double *result,
    void *params);
struct GeneralSpec {
    uint32_t K, M, W;
    const double *weights;   // host [num_samples x K]
    general_stat_func f;
    void *params;
    const double *windows;   // host [W + 1]
    uint32_t options;
    double *result;          // host [W x M]
};
int run_general_stat(const Plan *plan, const GeneralSpec &spec);
int run_sample_count_stat(const Plan *plan, const StatSpec &spec);
int run_weighted_stat(const Plan *plan, const StatSpec &spec);
int run_trees_at(const Plan *plan, uint64_t nq, const double *positions, const int32_t *tracked,
    uint64_t num_tracked, int32_t *out_parent, int32_t *out_count);

// matrix.cu
int run_genotype_matrix(const Plan *plan, const int32_t *samples, uint64_t num_samples,
    uint32_t options, int8_t *genotypes, uint64_t first_site, uint64_t num_sites);
int run_divergence_matrix(const Plan *plan, uint64_t nsets, const uint64_t *sizes, const int32_t *sets,
    uint64_t num_windows, const double *windows, uint32_t options, double *result);

}  // namespace tskb

// the opaque handle of the C ABI
struct tskb_treeseq {
    tskb::Plan *plan;
};
