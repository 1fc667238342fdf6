// api.cu -- the extern "C" boundary (include/tskit_b200.h): argument validation with the
// reference's error codes and precedence, then dispatch to the device engine.
// There is deliberately no host implementation of any statistic in this library.
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "plan.cuh"

using namespace tskb;

namespace {

template <typename Fn>
int guarded(Fn &&fn) {
    try {
        return fn();
    } catch (const CudaFail &f) {
        char buf[512];
        snprintf(buf, sizeof(buf), "%s: %s (%s:%d)", cudaGetErrorName(f.err), f.what, f.file, f.line);
        last_error_string() = buf;
        cudaGetLastError();
        return f.err == cudaErrorMemoryAllocation ? TSKB_ERR_NO_MEMORY : TSKB_ERR_CUDA;
    } catch (int code) {
        return code;
    } catch (const std::bad_alloc &) {
        return TSKB_ERR_NO_MEMORY;
    }
}

// tsk_treeseq_check_windows with TSK_REQUIRE_FULL_SPAN (trees.c:1244-1286)
int check_windows(const Plan &P, uint64_t num_windows, const double *windows, bool full_span) {
    if (num_windows < 1) return TSKB_ERR_BAD_NUM_WINDOWS;
    if (full_span) {
        if (windows[0] != 0) return TSKB_ERR_BAD_WINDOWS;
        if (windows[num_windows] != P.L) return TSKB_ERR_BAD_WINDOWS;
    } else {
        if (windows[0] < 0) return TSKB_ERR_BAD_WINDOWS;
        if (windows[num_windows] > P.L) return TSKB_ERR_BAD_WINDOWS;
    }
    for (uint64_t j = 0; j < num_windows; j++) {
        if (windows[j] >= windows[j + 1]) return TSKB_ERR_BAD_WINDOWS;
    }
    return 0;
}

// tsk_treeseq_check_sample_sets (trees.c:2114-2149) followed by the duplicate
// check of tsk_treeseq_sample_count_stat (trees.c:2195-2213)
int check_sample_sets(const Plan &P, uint64_t K, const uint64_t *sizes, const int32_t *sets) {
    if (K == 0) return TSKB_ERR_INSUFFICIENT_SAMPLE_SETS;
    // one pass; a duplicate is reported only if no set fails the earlier checks (the reference
    // validates every set before it looks for duplicates).  `seen` persists across calls: an entry
    // is current when it lies in this call's stamp range, so it is never cleared.
    std::lock_guard<std::mutex> lock(P.mu);
    std::vector<uint32_t> &seen = P.seen_stamp;
    if (seen.size() != P.num_samples || P.seen_epoch > 0xffffffffu - 2 * K - 2) {
        seen.assign(P.num_samples, 0);
        P.seen_epoch = 0;
    }
    const uint32_t base = (uint32_t) P.seen_epoch + 1;
    P.seen_epoch += K;
    const int32_t N = (int32_t) P.N;
    const int32_t *map = P.sample_index_map.data();
    bool duplicate = false;
    uint64_t j = 0;
    for (uint64_t k = 0; k < K; k++) {
        if (sizes[k] == 0) return TSKB_ERR_EMPTY_SAMPLE_SET;
        const uint32_t stamp = base + (uint32_t) k;
        for (uint64_t l = 0; l < sizes[k]; l++, j++) {
            const int32_t u = sets[j];
            if (u < 0 || u >= N) return TSKB_ERR_NODE_OUT_OF_BOUNDS;
            const int32_t si = map[u];
            if (si == -1) return TSKB_ERR_BAD_SAMPLES;
            duplicate |= seen[si] == stamp;
            seen[si] = stamp;
        }
    }
    return duplicate ? TSKB_ERR_DUPLICATE_SAMPLE : 0;
}

int tuple_width(int stat_id) {
    switch (stat_id) {
        case STAT_DIVERGENCE: case STAT_Y2: case STAT_F2: case STAT_RELATEDNESS:
        case STAT_RELATEDNESS_NC:
            return 2;
        case STAT_Y3: case STAT_F3:
            return 3;
        case STAT_F4:
            return 4;
    }
    return 0;
}

constexpr uint64_t MAX_STATE_DIM = 8;  // sample sets (state columns) of one sweep

// More sample sets than one sweep carries: the result columns are independent of each other (a
// column reads only the sets of its own index tuple), so they are computed in batches whose
// tuples touch at most MAX_STATE_DIM distinct sets, each batch one sweep over re-numbered sets.
// The centred genetic_relatedness reads every set in every column (the mean over all sets,
// trees.c:4899-4959) and cannot be split.
int weighted_stat(const tskb_treeseq_t *self, int stat_id, uint64_t cols, const std::vector<double> &W,
    uint64_t result_dim, int tw, const int32_t *tuples, uint64_t num_windows, const double *windows,
    uint32_t options, double *result, const double *table = nullptr, uint64_t table_rows = 0);

// The centred genetic_relatedness reads every set in every column: the mean over all K sets
// (trees.c:4729-4753).  Beyond one sweep's worth of sets that mean travels as ONE extra fp64 state
// column -- per sample, the sum of 1/n_k over the sets holding it -- next to the 0/1 indicator
// columns of the sets a batch of columns needs (counts are exact in fp64), through the weighted
// engine and its column batching.
int centred_relatedness_many_sets(const tskb_treeseq_t *self, uint64_t K, const uint64_t *sizes,
    const int32_t *sets, uint64_t M, const int32_t *tuples, uint64_t W, const double *windows,
    uint32_t options, double *result) {
    const Plan &P = *self->plan;
    const uint64_t n = P.num_samples, cols = K + 1;
    std::vector<double> Wt(n * cols, 0.0);
    uint64_t j = 0;
    for (uint64_t k = 0; k < K; k++) {
        const double inv = 1.0 / (double) sizes[k];
        for (uint64_t l = 0; l < sizes[k]; l++, j++) {
            const int32_t si = P.sample_index_map[sets[j]];  // validated by check_sample_sets
            Wt[(uint64_t) si * cols + k] = 1.0;
            Wt[(uint64_t) si * cols + K] += inv;
        }
    }
    return weighted_stat(self, STAT_REL_SIDE, cols, Wt, M, 2, tuples, W, windows, options, result, nullptr, K);
}

int batched_sample_count_stat(const Plan &P, int stat_id, int tw, uint64_t K, const uint64_t *sizes,
    const int32_t *sets, uint64_t M, const int32_t *tuples, uint64_t W, const double *windows,
    uint32_t options, double *result) {
    if (stat_id == STAT_RELATEDNESS) return TSKB_ERR_UNSUPPORTED;  // centred_relatedness_many_sets
    std::vector<uint64_t> off(K + 1, 0);
    for (uint64_t k = 0; k < K; k++) off[k + 1] = off[k] + sizes[k];
    const int width = tw > 0 ? tw : 1;
    std::vector<int32_t> local(K, -1), b_tuples, b_sets, used;
    std::vector<uint64_t> b_sizes, b_cols;
    std::vector<double> out;
    auto flush = [&]() -> int {
        if (b_cols.empty()) return 0;
        const uint64_t Mb = b_cols.size();
        out.assign(W * Mb, 0.0);
        StatSpec sp = {};
        sp.stat_id = stat_id;
        sp.K = (uint32_t) b_sizes.size();
        sp.M = (uint32_t) Mb;
        sp.tuple = (uint32_t) tw;
        sp.sizes = b_sizes.data();
        sp.sets = b_sets.data();
        sp.indexes = b_tuples.data();
        sp.W = (uint32_t) W;
        sp.windows = windows;
        sp.options = options;
        sp.result = out.data();
        int ret = run_sample_count_stat(&P, sp);
        if (ret != 0) return ret;
        for (uint64_t w = 0; w < W; w++) {
            for (uint64_t j = 0; j < Mb; j++) result[w * M + b_cols[j]] = out[w * Mb + j];
        }
        for (int32_t k : used) local[k] = -1;
        used.clear(); b_tuples.clear(); b_sets.clear(); b_sizes.clear(); b_cols.clear();
        return 0;
    };
    for (uint64_t m = 0; m < M; m++) {
        // sets of column m: its tuple, or set m itself for a one-way statistic
        int32_t need[4];
        for (int a = 0; a < width; a++) need[a] = tw > 0 ? tuples[m * tw + a] : (int32_t) m;
        uint64_t fresh = 0;
        for (int a = 0; a < width; a++) {
            bool seen = local[need[a]] >= 0;
            for (int c = 0; c < a; c++) seen |= need[c] == need[a];
            fresh += !seen;
        }
        if (b_sizes.size() + fresh > MAX_STATE_DIM) {
            int ret = flush();
            if (ret != 0) return ret;
        }
        for (int a = 0; a < width; a++) {
            const int32_t k = need[a];
            if (local[k] < 0) {
                local[k] = (int32_t) b_sizes.size();
                used.push_back(k);
                b_sizes.push_back(sizes[k]);
                b_sets.insert(b_sets.end(), sets + off[k], sets + off[k + 1]);
            }
            if (tw > 0) b_tuples.push_back(local[k]);
        }
        b_cols.push_back(m);
        // a one-way batch must keep column j = set j: flush when full
        if (tw == 0 && b_sizes.size() == MAX_STATE_DIM) {
            int ret = flush();
            if (ret != 0) return ret;
        }
    }
    return flush();
}

// Common path of every sample-count statistic; check precedence follows the
// reference call chain: index tuples (check_sample_stat_inputs, trees.c:4667)
// -> sample sets (trees.c:2191) -> duplicates (2207) -> mode (2053) -> dims
// (2058-2065) -> windows (2070) -> time units (1383).
int sample_count_stat(const tskb_treeseq_t *self, int stat_id, uint64_t K, const uint64_t *sizes,
    const int32_t *sets, bool sets_on_device, const int32_t *host_sets_for_checks,
    uint64_t num_tuples, const int32_t *tuples, uint64_t result_dim, uint64_t table_rows,
    const double *f_table, uint64_t num_windows, const double *windows, uint32_t options,
    double *result, bool result_on_device) {
    if (self == nullptr || self->plan == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    if (K > 0 && (sizes == nullptr || sets == nullptr)) return TSKB_ERR_BAD_PARAM_VALUE;
    if (result == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    const Plan &P = *self->plan;
    return guarded([&]() -> int {
        int tw = tuple_width(stat_id);
        if (stat_id == STAT_RELATEDNESS && (options & TSKB_STAT_NONCENTRED)) {
            stat_id = STAT_RELATEDNESS_NC;
        }
        uint64_t M = result_dim;
        if (tw > 0) {
            if (K < 1) return TSKB_ERR_INSUFFICIENT_SAMPLE_SETS;
            if (num_tuples < 1) return TSKB_ERR_INSUFFICIENT_INDEX_TUPLES;
            for (uint64_t j = 0; j < num_tuples * tw; j++) {
                if (tuples[j] < 0 || tuples[j] >= (int32_t) K) return TSKB_ERR_BAD_SAMPLE_SET_INDEX;
            }
            M = num_tuples;
        } else if (stat_id != STAT_TABULATED) {
            M = K;
        }
        // Sample sets are validated on the device while their weights are written (stats.cu:
        // k_set_weights) and the verdict is read back with the result.  The host loop runs only
        // where it decides the precedence of an error found below, or for an empty set (whose
        // position among element errors matters), or when there are too many sets for one sweep.
        bool sets_checked = false;
        auto host_check = [&]() -> int {
            if (sets_checked || (host_sets_for_checks == nullptr && K != 0)) return 0;
            sets_checked = true;
            return check_sample_sets(P, K, sizes, host_sets_for_checks);
        };
        bool any_empty = K == 0 || K > MAX_STATE_DIM;
        for (uint64_t k = 0; k < K; k++) any_empty |= sizes[k] == 0;
        if (any_empty) {
            int ret = host_check();
            if (ret != 0) return ret;
        }
#define LATER(code) do { int r__ = host_check(); return r__ != 0 ? r__ : (code); } while (0)
        bool site = options & TSKB_STAT_SITE, branch = options & TSKB_STAT_BRANCH,
             node = options & TSKB_STAT_NODE;
        if (!(site || branch || node)) {
            site = true;
            options |= TSKB_STAT_SITE;
        }
        if (site + branch + node > 1) LATER(TSKB_ERR_MULTIPLE_STAT_MODES);
        if (K < 1) LATER(TSKB_ERR_BAD_STATE_DIMS);
        if (M < 1) LATER(TSKB_ERR_BAD_RESULT_DIMS);
        double default_windows[2] = { 0, P.L };
        if (windows == nullptr) {
            num_windows = 1;
            windows = default_windows;
        } else {
            int ret = check_windows(P, num_windows, windows, true);
            if (ret != 0) LATER(ret);
        }
        // node mode needs the plan that keeps every piece (TSKB_INIT_NODE_MODE), a host result and
        // one sweep's worth of sample sets; W x N x M doubles must fit the device
        if (node && (!P.all_pieces || result_on_device || K > MAX_STATE_DIM
                        || (double) num_windows * (double) P.N * (double) M > 1e9)) {
            LATER(TSKB_ERR_UNSUPPORTED);
        }
        if (branch && P.time_uncalibrated && !(options & TSKB_STAT_ALLOW_TIME_UNCALIBRATED)) {
            LATER(TSKB_ERR_TIME_UNCALIBRATED);
        }
        if (stat_id == STAT_TABULATED && K != 1) LATER(TSKB_ERR_UNSUPPORTED);
        if (K > MAX_STATE_DIM) {
            if (sets_on_device || result_on_device) LATER(TSKB_ERR_UNSUPPORTED);
            if (stat_id == STAT_RELATEDNESS) {
                return centred_relatedness_many_sets(self, K, sizes, sets, M, tuples, num_windows, windows,
                    options, result);
            }
            return batched_sample_count_stat(P, stat_id, tw, K, sizes, sets, M, tuples, num_windows,
                windows, options, result);
        }
        StatSpec sp = {};
        sp.stat_id = stat_id;
        sp.K = (uint32_t) K;
        sp.M = (uint32_t) M;
        sp.tuple = (uint32_t) tw;
        sp.sizes = sizes;
        sp.sets = sets;
        sp.sets_on_device = sets_on_device;
        sp.indexes = tuples;
        sp.W = (uint32_t) num_windows;
        sp.windows = windows;
        sp.options = options;
        sp.result = result;
        sp.result_on_device = result_on_device;
        sp.f_table = f_table;
        sp.table_rows = table_rows;
        return run_sample_count_stat(&P, sp);
#undef LATER
    });
}

}  // namespace

extern "C" {

int tskb_treeseq_init(tskb_treeseq_t **self, const tskb_tables_t *tables, int device,
    double range_left, double range_right, uint32_t options) {
    if (self == nullptr || tables == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    *self = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        last_error_string() = "no CUDA device visible; this engine has no CPU fallback";
        return TSKB_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) return TSKB_ERR_BAD_PARAM_VALUE;
    if (!(range_left >= 0 && range_right <= tables->sequence_length && range_left < range_right)) {
        return TSKB_ERR_BAD_PARAM_VALUE;
    }
    if (tables->num_edges > 0
        && (tables->edge_insertion_order == nullptr || tables->edge_removal_order == nullptr)) {
        return TSKB_ERR_BAD_PARAM_VALUE;
    }
    return guarded([&]() -> int {
        Plan *p = build_plan(tables, device, range_left, range_right, options);
        *self = new tskb_treeseq{ p };
        return 0;
    });
}

int tskb_treeseq_free(tskb_treeseq_t *self) {
    if (self != nullptr) {
        if (self->plan != nullptr) {
            cudaSetDevice(self->plan->device);
            delete self->plan;
        }
        delete self;
    }
    return 0;
}

const char *tskb_last_cuda_error(void) { return last_error_string().c_str(); }

const char *tskb_strerror(int err) {
    // messages of tsk_strerror (c/tskit/core.c) for the codes this path can return
    switch (err) {
        case 0: return "Normal exit condition. This is not an error!";
        case TSKB_ERR_NO_MEMORY: return "Out of memory. (TSK_ERR_NO_MEMORY)";
        case TSKB_ERR_BAD_PARAM_VALUE: return "Bad parameter value provided. (TSK_ERR_BAD_PARAM_VALUE)";
        case TSKB_ERR_NODE_OUT_OF_BOUNDS: return "Node out of bounds. (TSK_ERR_NODE_OUT_OF_BOUNDS)";
        case TSKB_ERR_BAD_OFFSET: return "Bad offset provided in input array. (TSK_ERR_BAD_OFFSET)";
        case TSKB_ERR_EDGE_OUT_OF_BOUNDS: return "Edge out of bounds. (TSK_ERR_EDGE_OUT_OF_BOUNDS)";
        case TSKB_ERR_SITE_OUT_OF_BOUNDS: return "Site out of bounds. (TSK_ERR_SITE_OUT_OF_BOUNDS)";
        case TSKB_ERR_MUTATION_OUT_OF_BOUNDS: return "Mutation out of bounds. (TSK_ERR_MUTATION_OUT_OF_BOUNDS)";
        case TSKB_ERR_TIME_NONFINITE: return "Times must be finite. (TSK_ERR_TIME_NONFINITE)";
        case TSKB_ERR_GENOME_COORDS_NONFINITE: return "Genome coordinates must be finite numbers. (TSK_ERR_GENOME_COORDS_NONFINITE)";
        case TSKB_ERR_NULL_PARENT: return "Edge parent is null. (TSK_ERR_NULL_PARENT)";
        case TSKB_ERR_NULL_CHILD: return "Edge child is null. (TSK_ERR_NULL_CHILD)";
        case TSKB_ERR_BAD_NODE_TIME_ORDERING: return "time[parent] must be greater than time[child]. (TSK_ERR_BAD_NODE_TIME_ORDERING)";
        case TSKB_ERR_BAD_EDGE_INTERVAL: return "Bad edge interval where right <= left. (TSK_ERR_BAD_EDGE_INTERVAL)";
        case TSKB_ERR_RIGHT_GREATER_SEQ_LENGTH: return "Right coordinate > sequence length. (TSK_ERR_RIGHT_GREATER_SEQ_LENGTH)";
        case TSKB_ERR_LEFT_LESS_ZERO: return "Left coordinate must be >= 0. (TSK_ERR_LEFT_LESS_ZERO)";
        case TSKB_ERR_UNSORTED_SITES: return "Sites must be provided in strictly increasing position order. (TSK_ERR_UNSORTED_SITES)";
        case TSKB_ERR_DUPLICATE_SITE_POSITION: return "Duplicate site positions. (TSK_ERR_DUPLICATE_SITE_POSITION)";
        case TSKB_ERR_BAD_SITE_POSITION: return "Site positions must be between 0 and sequence_length. (TSK_ERR_BAD_SITE_POSITION)";
        case TSKB_ERR_MUTATION_PARENT_DIFFERENT_SITE: return "Specified parent mutation is at a different site. (TSK_ERR_MUTATION_PARENT_DIFFERENT_SITE)";
        case TSKB_ERR_MUTATION_PARENT_EQUAL: return "Parent mutation refers to itself. (TSK_ERR_MUTATION_PARENT_EQUAL)";
        case TSKB_ERR_MUTATION_PARENT_AFTER_CHILD: return "Parent mutation ID must be < current ID. (TSK_ERR_MUTATION_PARENT_AFTER_CHILD)";
        case TSKB_ERR_UNSORTED_MUTATIONS: return "Mutations must be provided in non-decreasing site order and non-increasing time order within each site. (TSK_ERR_UNSORTED_MUTATIONS)";
        case TSKB_ERR_BAD_SEQUENCE_LENGTH: return "Sequence length must be > 0. (TSK_ERR_BAD_SEQUENCE_LENGTH)";
        case TSKB_ERR_DUPLICATE_SAMPLE: return "Duplicate sample value. (TSK_ERR_DUPLICATE_SAMPLE)";
        case TSKB_ERR_BAD_SAMPLES: return "The nodes provided are not samples. (TSK_ERR_BAD_SAMPLES)";
        case TSKB_ERR_BAD_NUM_WINDOWS: return "Must have at least one window, [0, L]. (TSK_ERR_BAD_NUM_WINDOWS)";
        case TSKB_ERR_BAD_WINDOWS: return "Windows must be increasing list [0, ..., L]. (TSK_ERR_BAD_WINDOWS)";
        case TSKB_ERR_MULTIPLE_STAT_MODES: return "Cannot specify more than one stats mode. (TSK_ERR_MULTIPLE_STAT_MODES)";
        case TSKB_ERR_BAD_STATE_DIMS: return "Must have state dimension >= 1. (TSK_ERR_BAD_STATE_DIMS)";
        case TSKB_ERR_BAD_RESULT_DIMS: return "Must have result dimension >= 1. (TSK_ERR_BAD_RESULT_DIMS)";
        case TSKB_ERR_INSUFFICIENT_SAMPLE_SETS: return "Insufficient sample sets provided. (TSK_ERR_INSUFFICIENT_SAMPLE_SETS)";
        case TSKB_ERR_INSUFFICIENT_INDEX_TUPLES: return "Insufficient sample set index tuples provided. (TSK_ERR_INSUFFICIENT_INDEX_TUPLES)";
        case TSKB_ERR_BAD_SAMPLE_SET_INDEX: return "Sample set index out of bounds. (TSK_ERR_BAD_SAMPLE_SET_INDEX)";
        case TSKB_ERR_EMPTY_SAMPLE_SET: return "Samples cannot be empty. (TSK_ERR_EMPTY_SAMPLE_SET)";
        case TSKB_ERR_UNSUPPORTED_STAT_MODE: return "Requested statistics mode not supported for this method. (TSK_ERR_UNSUPPORTED_STAT_MODE)";
        case TSKB_ERR_TIME_UNCALIBRATED: return "Statistics using branch lengths cannot be calculated when time_units is 'uncalibrated'. (TSK_ERR_TIME_UNCALIBRATED)";
        case TSKB_ERR_STAT_POLARISED_UNSUPPORTED: return "The TSK_STAT_POLARISED option is not supported by this statistic. (TSK_ERR_STAT_POLARISED_UNSUPPORTED)";
        case TSKB_ERR_BAD_TIME_WINDOWS_DIM: return "Must have at least one time window. (TSK_ERR_BAD_TIME_WINDOWS_DIM)";
        case TSKB_ERR_BAD_TIME_WINDOWS: return "Time windows must start at zero and be strictly increasing. (TSK_ERR_BAD_TIME_WINDOWS)";
        case TSKB_ERR_INSUFFICIENT_WEIGHTS: return "Insufficient weights provided (at least 1 required). (TSK_ERR_INSUFFICIENT_WEIGHTS)";
        case TSKB_ERR_CUDA: return "CUDA runtime error (see tskb_last_cuda_error)";
        case TSKB_ERR_BAD_INDEX_ORDER: return "Edge indexes are not in the order tsk_table_collection_build_index produces";
        case TSKB_ERR_UNSUPPORTED: return "Valid tskit call that the B200 engine does not accelerate";
        case TSKB_ERR_NO_DEVICE: return "No CUDA device: the B200 engine has no CPU fallback";
    }
    return "Unknown error";
}

#define ONE_WAY(NAME, ID)                                                                       \
    int tskb_treeseq_##NAME(const tskb_treeseq_t *self, uint64_t num_sample_sets,                \
        const uint64_t *sample_set_sizes, const int32_t *sample_sets, uint64_t num_windows,     \
        const double *windows, uint32_t options, double *result) {                              \
        return sample_count_stat(self, ID, num_sample_sets, sample_set_sizes, sample_sets,      \
            false, sample_sets, 0, nullptr, num_sample_sets, 0, nullptr, num_windows, windows,  \
            options, result, false);                                                            \
    }
ONE_WAY(diversity, STAT_DIVERSITY)
ONE_WAY(segregating_sites, STAT_SEGSITES)
ONE_WAY(Y1, STAT_Y1)

#define K_WAY(NAME, ID)                                                                         \
    int tskb_treeseq_##NAME(const tskb_treeseq_t *self, uint64_t num_sample_sets,                \
        const uint64_t *sample_set_sizes, const int32_t *sample_sets,                           \
        uint64_t num_index_tuples, const int32_t *index_tuples, uint64_t num_windows,           \
        const double *windows, uint32_t options, double *result) {                              \
        return sample_count_stat(self, ID, num_sample_sets, sample_set_sizes, sample_sets,      \
            false, sample_sets, num_index_tuples, index_tuples, num_index_tuples, 0, nullptr,   \
            num_windows, windows, options, result, false);                                      \
    }
K_WAY(divergence, STAT_DIVERGENCE)
K_WAY(Y2, STAT_Y2)
K_WAY(f2, STAT_F2)
K_WAY(genetic_relatedness, STAT_RELATEDNESS)
K_WAY(Y3, STAT_Y3)
K_WAY(f3, STAT_F3)
K_WAY(f4, STAT_F4)

}  // extern "C" (reopened below)

namespace {

// Common path of the weighted statistics: `cols` state columns of pre-processed weights (what
// the reference hands to tsk_treeseq_general_stat); checks in general_stat's order
// (trees.c:2035-2095): mode -> dims -> windows -> time units.
int weighted_stat(const tskb_treeseq_t *self, int stat_id, uint64_t cols, const std::vector<double> &W,
    uint64_t result_dim, int tw, const int32_t *tuples, uint64_t num_windows, const double *windows,
    uint32_t options, double *result, const double *table, uint64_t table_rows) {
    const Plan &P = *self->plan;
    return guarded([&]() -> int {
        bool site = options & TSKB_STAT_SITE, branch = options & TSKB_STAT_BRANCH,
             node = options & TSKB_STAT_NODE;
        if (!(site || branch || node)) {
            site = true;
            options |= TSKB_STAT_SITE;
        }
        if (site + branch + node > 1) return TSKB_ERR_MULTIPLE_STAT_MODES;
        if (cols < 1) return TSKB_ERR_BAD_STATE_DIMS;
        if (result_dim < 1) return TSKB_ERR_BAD_RESULT_DIMS;
        double default_windows[2] = { 0, P.L };
        if (windows == nullptr) {
            num_windows = 1;
            windows = default_windows;
        } else {
            int ret = check_windows(P, num_windows, windows, true);
            if (ret != 0) return ret;
        }
        if (node && (!P.all_pieces || cols > MAX_STATE_DIM
                        || (double) num_windows * (double) P.N * (double) result_dim > 1e9)) {
            return TSKB_ERR_UNSUPPORTED;
        }
        if (branch && P.time_uncalibrated && !(options & TSKB_STAT_ALLOW_TIME_UNCALIBRATED)) {
            return TSKB_ERR_TIME_UNCALIBRATED;
        }
        if (cols > MAX_STATE_DIM && stat_id == STAT_TRAIT_LM) return TSKB_ERR_UNSUPPORTED;
        if (cols > MAX_STATE_DIM) {
            // Result columns are independent: batches whose columns read at most MAX_STATE_DIM state
            // columns (the frequency column, where there is one, travels with every batch, last).
            const bool freq = stat_id != STAT_TRAIT_COV;
            const uint64_t n = P.num_samples, room = MAX_STATE_DIM - (freq ? 1 : 0);
            const int width = tw > 0 ? tw : 1;
            std::vector<int32_t> local(cols, -1), used, b_tuples;
            std::vector<uint64_t> b_cols;
            auto flush = [&]() -> int {
                if (b_cols.empty()) return 0;
                const uint64_t kb = used.size() + (freq ? 1 : 0), Mb = b_cols.size();
                std::vector<double> Wb(n * kb), out(num_windows * Mb, 0.0);
                for (uint64_t j = 0; j < n; j++) {
                    for (uint64_t q = 0; q < used.size(); q++) Wb[j * kb + q] = W[j * cols + used[q]];
                    if (freq) Wb[j * kb + kb - 1] = W[j * cols + cols - 1];
                }
                for (auto &x : b_tuples) {
                    if (x < 0) x = (int32_t) kb - 1;
                }
                int ret = weighted_stat(self, stat_id, kb, Wb, Mb, tw, tw > 0 ? b_tuples.data() : nullptr,
                    num_windows, windows, options, out.data(), table, table_rows);
                if (ret != 0) return ret;
                for (uint64_t w = 0; w < num_windows; w++) {
                    for (uint64_t q = 0; q < Mb; q++) result[w * result_dim + b_cols[q]] = out[w * Mb + q];
                }
                for (int32_t k : used) local[k] = -1;
                used.clear(); b_tuples.clear(); b_cols.clear();
                return 0;
            };
            for (uint64_t m = 0; m < result_dim; m++) {
                int32_t need[2];
                for (int a = 0; a < width; a++) need[a] = tw > 0 ? tuples[m * tw + a] : (int32_t) m;
                uint64_t fresh = 0;
                for (int a = 0; a < width; a++) {
                    bool seen = (freq && need[a] == (int32_t) cols - 1) || local[need[a]] >= 0;
                    for (int c = 0; c < a; c++) seen |= need[c] == need[a];
                    fresh += !seen;
                }
                if (used.size() + fresh > room) {
                    int ret = flush();
                    if (ret != 0) return ret;
                }
                for (int a = 0; a < width; a++) {
                    const int32_t k = need[a];
                    const bool is_freq = freq && k == (int32_t) cols - 1;
                    if (!is_freq && local[k] < 0) {
                        local[k] = (int32_t) used.size();
                        used.push_back(k);
                    }
                    // the frequency column's local index is only known at flush time: mark it
                    if (tw > 0) b_tuples.push_back(is_freq ? -1 : local[k]);
                }
                b_cols.push_back(m);
                // one-way statistics keep result column q = state column q: flush when full
                if (tw == 0 && used.size() == room) {
                    int ret = flush();
                    if (ret != 0) return ret;
                }
            }
            return flush();
        }
        // total_weight of general_stat: summed over the samples in order (trees.c:2003-2010)
        std::vector<double> totals(cols, 0.0);
        for (uint64_t j = 0; j < P.num_samples; j++) {
            for (uint64_t k = 0; k < cols; k++) totals[k] += W[j * cols + k];
        }
        StatSpec sp = {};
        sp.stat_id = stat_id;
        sp.K = (uint32_t) cols;
        sp.M = (uint32_t) result_dim;
        sp.tuple = (uint32_t) tw;
        sp.indexes = tuples;
        sp.W = (uint32_t) num_windows;
        sp.windows = windows;
        sp.options = options;
        sp.result = result;
        sp.weights = W.data();
        sp.column_totals = totals.data();
        sp.f_table = table;
        sp.table_rows = table_rows;
        return run_weighted_stat(&P, sp);
    });
}

}  // namespace

extern "C" {

/* tsk_treeseq_trait_covariance (trees.c:3976-4022): centre the weights, state = their sums */
int tskb_treeseq_trait_covariance(const tskb_treeseq_t *self, uint64_t num_weights, const double *weights,
    uint64_t num_windows, const double *windows, uint32_t options, double *result) {
    if (self == nullptr || self->plan == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    if (num_weights == 0) return TSKB_ERR_INSUFFICIENT_WEIGHTS;
    const uint64_t n = self->plan->num_samples, K = num_weights;
    std::vector<double> means(K, 0.0), W(n * K);
    for (uint64_t j = 0; j < n; j++) {
        for (uint64_t k = 0; k < K; k++) means[k] += weights[j * K + k];
    }
    for (uint64_t k = 0; k < K; k++) means[k] /= (double) n;
    for (uint64_t j = 0; j < n; j++) {
        for (uint64_t k = 0; k < K; k++) W[j * K + k] = weights[j * K + k] - means[k];
    }
    return weighted_stat(self, STAT_TRAIT_COV, K, W, K, 0, nullptr, num_windows, windows, options, result);
}

/* tsk_treeseq_trait_correlation (trees.c:4051-4110): standardise the weights, append 1/n */
int tskb_treeseq_trait_correlation(const tskb_treeseq_t *self, uint64_t num_weights, const double *weights,
    uint64_t num_windows, const double *windows, uint32_t options, double *result) {
    if (self == nullptr || self->plan == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    if (num_weights < 1) return TSKB_ERR_INSUFFICIENT_WEIGHTS;
    const uint64_t n = self->plan->num_samples, K = num_weights;
    std::vector<double> means(K, 0.0), meansqs(K, 0.0), sds(K, 0.0), W(n * (K + 1));
    for (uint64_t j = 0; j < n; j++) {
        for (uint64_t k = 0; k < K; k++) {
            means[k] += weights[j * K + k];
            meansqs[k] += weights[j * K + k] * weights[j * K + k];
        }
    }
    for (uint64_t k = 0; k < K; k++) {
        means[k] /= (double) n;
        meansqs[k] -= means[k] * means[k] * (double) n;
        meansqs[k] /= (double) (n - 1);
        sds[k] = sqrt(meansqs[k]);
    }
    for (uint64_t j = 0; j < n; j++) {
        for (uint64_t k = 0; k < K; k++) W[j * (K + 1) + k] = (weights[j * K + k] - means[k]) / sds[k];
        W[j * (K + 1) + K] = 1.0 / (double) n;  // frequency column
    }
    return weighted_stat(self, STAT_TRAIT_CORR, K + 1, W, K, 0, nullptr, num_windows, windows, options,
        result);
}

/* tsk_treeseq_trait_linear_model (trees.c:4153-4219): covariates already orthonormalised by the
 * caller (as the reference assumes); state = traits | covariates | number of samples below */
int tskb_treeseq_trait_linear_model(const tskb_treeseq_t *self, uint64_t num_weights, const double *weights,
    uint64_t num_covariates, const double *covariates, uint64_t num_windows, const double *windows,
    uint32_t options, double *result) {
    if (self == nullptr || self->plan == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    if (num_weights < 1) return TSKB_ERR_INSUFFICIENT_WEIGHTS;
    const uint64_t n = self->plan->num_samples, K = num_weights, C = num_covariates, cols = K + C + 1;
    std::vector<double> V(std::max<uint64_t>(C * K, 1), 0.0), W(n * cols);
    for (uint64_t k = 0; k < n; k++) {  // V = weights^T covariates, in the reference's loop order
        for (uint64_t i = 0; i < K; i++) {
            for (uint64_t j = 0; j < C; j++) V[i * C + j] += weights[k * K + i] * covariates[k * C + j];
        }
    }
    for (uint64_t k = 0; k < n; k++) {
        for (uint64_t i = 0; i < K; i++) W[k * cols + i] = weights[k * K + i];
        for (uint64_t i = 0; i < C; i++) W[k * cols + K + i] = covariates[k * C + i];
        W[k * cols + K + C] = 1.0;
    }
    return weighted_stat(self, STAT_TRAIT_LM, cols, W, K, 0, nullptr, num_windows, windows, options, result,
        V.data(), C);
}

/* tsk_treeseq_genetic_relatedness_weighted (trees.c:4840-4897): append the 1/n column */
int tskb_treeseq_genetic_relatedness_weighted(const tskb_treeseq_t *self, uint64_t num_weights,
    const double *weights, uint64_t num_index_tuples, const int32_t *index_tuples, uint64_t num_windows,
    const double *windows, double *result, uint32_t options) {
    if (self == nullptr || self->plan == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    if (num_weights == 0) return TSKB_ERR_INSUFFICIENT_WEIGHTS;
    const uint64_t n = self->plan->num_samples, K = num_weights;
    for (uint64_t j = 0; j < 2 * num_index_tuples; j++) {
        // the reference indexes the state row unchecked; out of range is refused here
        if (index_tuples[j] < 0 || index_tuples[j] > (int32_t) K) return TSKB_ERR_BAD_SAMPLE_SET_INDEX;
    }
    std::vector<double> W(n * (K + 1));
    for (uint64_t j = 0; j < n; j++) {
        for (uint64_t k = 0; k < K; k++) W[j * (K + 1) + k] = weights[j * K + k];
        W[j * (K + 1) + K] = 1.0 / (double) n;
    }
    const int stat = (options & TSKB_STAT_NONCENTRED) ? STAT_REL_WEIGHTED_NC : STAT_REL_WEIGHTED;
    return weighted_stat(self, stat, K + 1, W, num_index_tuples, 2, index_tuples, num_windows, windows,
        options, result);
}

/* tsk_treeseq_genetic_relatedness_vector (trees.c:10772-10816): branch mode only; checks in the
 * reference's order (mode -> windows, which need not span the sequence -> focal nodes).  Weights are
 * centred before and the output rows after the device pass exactly as tsk_matvec_calculator_init /
 * _write_output do (trees.c:10512-10535, 10692-10713); span normalisation as trees.c:1920-1934. */
int tskb_treeseq_genetic_relatedness_vector(const tskb_treeseq_t *self, uint64_t num_weights,
    const double *weights, uint64_t num_windows, const double *windows, uint64_t num_focal_nodes,
    const int32_t *focal_nodes, double *result, uint32_t options) {
    if (self == nullptr || self->plan == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    if ((weights == nullptr && num_weights > 0) || (focal_nodes == nullptr && num_focal_nodes > 0)
        || result == nullptr) {
        return TSKB_ERR_BAD_PARAM_VALUE;
    }
    const Plan &P = *self->plan;
    return guarded([&]() -> int {
        const auto t0 = std::chrono::steady_clock::now();
        double ms_device_calls = 0;
        if (options & (TSKB_STAT_SITE | TSKB_STAT_NODE)) return TSKB_ERR_UNSUPPORTED_STAT_MODE;
        if (windows == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
        int ret = check_windows(P, num_windows, windows, false);
        if (ret != 0) return ret;
        const uint64_t n = P.num_samples, K = num_weights, nf = num_focal_nodes;
        bool needs_nodes = false;
        for (uint64_t j = 0; j < nf; j++) {
            if (focal_nodes[j] < 0 || (uint64_t) focal_nodes[j] >= P.N) return TSKB_ERR_NODE_OUT_OF_BOUNDS;
            needs_nodes |= P.sample_index_map[focal_nodes[j]] < 0;
        }
        if (needs_nodes && !P.all_pieces) return TSKB_ERR_UNSUPPORTED;  // needs TSKB_INIT_NODE_MODE
        const uint64_t row = nf * K;
        if ((double) num_windows * (double) row > 2e9 || row > 0xffffffffull) return TSKB_ERR_UNSUPPORTED;
        if (row == 0) return 0;  // nothing to write (the engine writes every entry otherwise)
        std::vector<double> means(K, 0.0);
        if (!(options & TSKB_STAT_NONCENTRED)) {
            for (uint64_t j = 0; j < n; j++) {
                for (uint64_t k = 0; k < K; k++) means[k] += weights[j * K + k];
            }
            for (uint64_t k = 0; k < K; k++) means[k] /= (double) n;
        }
        // state columns in batches of one sweep's width; a single batch reads the caller's weights
        // (when they are not centred) and writes the caller's result directly
        const bool centred = !(options & TSKB_STAT_NONCENTRED);
        const bool one_batch = K <= MAX_STATE_DIM;
        std::vector<double> Wb, out;
        for (uint64_t k0 = 0; k0 < K; k0 += MAX_STATE_DIM) {
            const uint64_t kb = K - k0 < MAX_STATE_DIM ? K - k0 : MAX_STATE_DIM;
            std::vector<double> totals(kb, 0.0);
            const double *w_in = weights;
            if (centred || !one_batch) {
                Wb.resize(n * kb);
                for (uint64_t j = 0; j < n; j++) {
                    for (uint64_t k = 0; k < kb; k++) Wb[j * kb + k] = weights[j * K + k0 + k] - means[k0 + k];
                }
                w_in = Wb.data();
            }
            if (!one_batch) out.resize(num_windows * nf * kb);
            StatSpec sp = {};
            sp.stat_id = STAT_REL_VECTOR;
            sp.K = (uint32_t) kb;
            sp.M = (uint32_t) (nf * kb);
            sp.W = (uint32_t) num_windows;
            sp.windows = windows;
            sp.options = TSKB_STAT_BRANCH;
            sp.result = one_batch ? result : out.data();
            sp.weights = w_in;
            sp.column_totals = totals.data();
            sp.focal = focal_nodes;
            sp.num_focal = nf;
            sp.focal_needs_nodes = needs_nodes;
            const auto t1 = std::chrono::steady_clock::now();
            ret = run_weighted_stat(&P, sp);
            ms_device_calls += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
            if (ret != 0) return ret;
            if (one_batch) break;
            for (uint64_t w = 0; w < num_windows; w++) {
                for (uint64_t j = 0; j < nf; j++) {
                    for (uint64_t k = 0; k < kb; k++) {
                        result[(w * nf + j) * K + k0 + k] = out[(w * nf + j) * kb + k];
                    }
                }
            }
        }
        for (uint64_t w = 0; w < num_windows; w++) {
            double *y = result + w * row;
            if (!(options & TSKB_STAT_NONCENTRED)) {
                std::vector<double> out_means(K, 0.0);
                for (uint64_t j = 0; j < nf; j++) {
                    for (uint64_t k = 0; k < K; k++) out_means[k] += y[j * K + k];
                }
                for (uint64_t k = 0; k < K; k++) out_means[k] /= (double) nf;
                for (uint64_t j = 0; j < nf; j++) {
                    for (uint64_t k = 0; k < K; k++) y[j * K + k] -= out_means[k];
                }
            }
            if (options & TSKB_STAT_SPAN_NORMALISE) {
                const double span = windows[w + 1] - windows[w];
                for (uint64_t i = 0; i < row; i++) y[i] /= span;
            }
        }
        if (getenv("TSKB_TIMING") != nullptr) {
            fprintf(stderr, "tskb timing: relatedness_vector total %.3f ms, of which engine calls %.3f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), ms_device_calls);
        }
        return 0;
    });
}

/* tsk_treeseq_allele_frequency_spectrum (trees.c:3814-3928), site mode; checks in its order */
int tskb_treeseq_allele_frequency_spectrum(const tskb_treeseq_t *self, uint64_t num_sample_sets,
    const uint64_t *sample_set_sizes, const int32_t *sample_sets, uint64_t num_windows, const double *windows,
    uint64_t num_time_windows, const double *time_windows, uint32_t options, double *result) {
    if (self == nullptr || self->plan == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    const Plan &P = *self->plan;
    return guarded([&]() -> int {
        bool site = options & TSKB_STAT_SITE, branch = options & TSKB_STAT_BRANCH;
        if (options & TSKB_STAT_NODE) return TSKB_ERR_UNSUPPORTED_STAT_MODE;
        if (!(site || branch)) {
            site = true;
            options |= TSKB_STAT_SITE;
        }
        if (site + branch > 1) return TSKB_ERR_MULTIPLE_STAT_MODES;
        double default_windows[2] = { 0, P.L };
        if (windows == nullptr) {
            num_windows = 1;
            windows = default_windows;
        } else {
            int ret = check_windows(P, num_windows, windows, true);
            if (ret != 0) return ret;
        }
        if (time_windows != nullptr) {  // tsk_treeseq_check_time_windows (trees.c:1288-1313)
            if (num_time_windows < 1) return TSKB_ERR_BAD_TIME_WINDOWS_DIM;
            if (time_windows[0] != 0.0) return TSKB_ERR_BAD_TIME_WINDOWS;
            for (uint64_t j = 0; j < num_time_windows; j++) {
                if (time_windows[j] >= time_windows[j + 1]) return TSKB_ERR_BAD_TIME_WINDOWS;
            }
            if (site && !(time_windows[0] == 0.0 && std::isinf((float) time_windows[1]))) {
                return TSKB_ERR_UNSUPPORTED_STAT_MODE;  // site mode has no time windows (trees.c:3868)
            }
        }
        int ret = check_sample_sets(P, num_sample_sets, sample_set_sizes, sample_sets);
        if (ret != 0) return ret;
        if (branch && P.time_uncalibrated && !(options & TSKB_STAT_ALLOW_TIME_UNCALIBRATED)) {
            return TSKB_ERR_TIME_UNCALIBRATED;  // trees.c:3727
        }
        // the default time window with node times >= 0: the branch length inside it is the whole branch;
        // any other time windows split every branch by time, which needs the node of every piece
        // (TSKB_INIT_NODE_MODE plans keep it; lowlevel.py stages one on first use)
        const bool default_tw = time_windows == nullptr
                                || (num_time_windows == 1 && time_windows[0] == 0.0 && std::isinf(time_windows[1]));
        const bool by_time = branch && (!default_tw || P.has_negative_time);
        if (by_time && !P.all_pieces) return TSKB_ERR_UNSUPPORTED;
        const double default_tws[2] = { 0.0, INFINITY };
        if (time_windows == nullptr) {
            num_time_windows = 1;
            time_windows = default_tws;
        }
        uint64_t total = 0, afs_size = 1;
        std::vector<uint32_t> dims;
        for (uint64_t k = 0; k < num_sample_sets; k++) {
            total += sample_set_sizes[k];
            afs_size *= sample_set_sizes[k] + 1;
            dims.push_back((uint32_t) sample_set_sizes[k] + 1);
            if (afs_size * num_windows * num_time_windows > (uint64_t) 2e9) return TSKB_ERR_UNSUPPORTED;
        }
        StatSpec sp = {};
        sp.stat_id = STAT_AFS;
        sp.M = 1;
        sp.W = (uint32_t) num_windows;
        sp.windows = windows;
        sp.options = options;
        sp.result = result;
        sp.afs_size = afs_size;
        sp.num_time_windows = (uint32_t) num_time_windows;
        sp.time_windows = by_time ? time_windows : nullptr;
        if (num_sample_sets + 1 > MAX_STATE_DIM) {
            // more than 7 sets: the spectrum's row-major coordinate is linear in the per-set counts, so it
            // travels as one fp64 state column (exact), next to the all-samples count
            if (num_sample_sets > 64 || afs_size >= (uint64_t(1) << 52)) return TSKB_ERR_UNSUPPORTED;
            const uint64_t n = P.num_samples;
            std::vector<double> Wt(n * 2, 0.0), totals(2, 0.0);
            std::vector<double> stride(num_sample_sets, 1.0);
            for (uint64_t k = num_sample_sets; k-- > 1;) stride[k - 1] = stride[k] * (double) dims[k];
            uint64_t j = 0;
            for (uint64_t k = 0; k < num_sample_sets; k++) {
                for (uint64_t l = 0; l < sample_set_sizes[k]; l++, j++) {
                    Wt[(uint64_t) P.sample_index_map[sample_sets[j]] * 2] += stride[k];
                }
            }
            for (uint64_t q = 0; q < n; q++) {
                Wt[q * 2 + 1] = 1.0;
                totals[0] += Wt[q * 2];
                totals[1] += 1.0;
            }
            sp.K = 2;
            sp.weights = Wt.data();
            sp.column_totals = totals.data();
            sp.afs_dims = dims.data();
            sp.afs_nsets = (uint32_t) num_sample_sets;
            return run_weighted_stat(&P, sp);
        }
        // state columns: the sets, then all samples (trees.c:3890-3910)
        std::vector<uint64_t> sizes(sample_set_sizes, sample_set_sizes + num_sample_sets);
        std::vector<int32_t> sets(sample_sets, sample_sets + total);
        sets.insert(sets.end(), P.samples.begin(), P.samples.end());
        sizes.push_back(P.num_samples);
        sp.K = (uint32_t) sizes.size();
        sp.sizes = sizes.data();
        sp.sets = sets.data();
        return run_sample_count_stat(&P, sp);
    });
}

int tskb_treeseq_sample_count_stat_tabulated(const tskb_treeseq_t *self,
    uint64_t num_sample_sets, const uint64_t *sample_set_sizes, const int32_t *sample_sets,
    uint64_t result_dim, uint64_t table_rows, const double *f_table, uint64_t num_windows,
    const double *windows, uint32_t options, double *result) {
    if (f_table == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    // the table must cover every count of the set; the sample-set arguments themselves are checked
    // (null pointers, K, empty sets ...) by sample_count_stat in the reference's precedence
    if (num_sample_sets == 1 && sample_set_sizes != nullptr && table_rows < sample_set_sizes[0] + 1) {
        return TSKB_ERR_BAD_PARAM_VALUE;
    }
    return sample_count_stat(self, STAT_TABULATED, num_sample_sets, sample_set_sizes,
        sample_sets, false, sample_sets, 0, nullptr, result_dim, table_rows, f_table,
        num_windows, windows, options, result, false);
}

int tskb_treeseq_general_stat(const tskb_treeseq_t *self, uint64_t state_dim, const double *weights,
    uint64_t result_dim, tskb_general_stat_func_t *f, void *f_params, uint64_t num_windows,
    const double *windows, uint32_t options, double *result) {
    if (self == nullptr || self->plan == nullptr || f == nullptr || result == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    if (state_dim > 0 && weights == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    const Plan &P = *self->plan;
    return guarded([&]() -> int {
        // checks in tsk_treeseq_general_stat's order (trees.c:2035-2095)
        bool site = options & TSKB_STAT_SITE, branch = options & TSKB_STAT_BRANCH, node = options & TSKB_STAT_NODE;
        if (!(site || branch || node)) {
            site = true;
            options |= TSKB_STAT_SITE;
        }
        if (site + branch + node > 1) return TSKB_ERR_MULTIPLE_STAT_MODES;
        if (state_dim < 1) return TSKB_ERR_BAD_STATE_DIMS;
        if (result_dim < 1) return TSKB_ERR_BAD_RESULT_DIMS;
        double default_windows[2] = { 0, P.L };
        if (windows == nullptr) {
            num_windows = 1;
            windows = default_windows;
        } else {
            int ret = check_windows(P, num_windows, windows, true);
            if (ret != 0) return ret;
        }
        if (branch && P.time_uncalibrated && !(options & TSKB_STAT_ALLOW_TIME_UNCALIBRATED)) {
            return TSKB_ERR_TIME_UNCALIBRATED;
        }
        if (node || state_dim > MAX_STATE_DIM) return TSKB_ERR_UNSUPPORTED;
        // Branch mode: the reference's running sum also takes 0 x f(state) from nodes without a branch
        // above them (trees.c:1339-1350), which matters exactly when f is NaN / inf there -- and whether it
        // is cannot be known without evaluating f at those states.  So the callback form runs on the
        // plan that keeps every piece (TSKB_INIT_NODE_MODE).
        if (branch && !P.all_pieces) return TSKB_ERR_UNSUPPORTED;
        GeneralSpec g = {};
        g.K = (uint32_t) state_dim; g.M = (uint32_t) result_dim; g.W = (uint32_t) num_windows;
        g.weights = weights; g.f = (general_stat_func) f; g.params = f_params;
        g.windows = windows; g.options = options; g.result = result;
        return run_general_stat(&P, g);
    });
}

int tskb_treeseq_stat_device(const tskb_treeseq_t *self, int stat_id, uint64_t num_sample_sets,
    const uint64_t *sample_set_sizes, const int32_t *d_sample_sets, uint64_t num_index_tuples,
    const int32_t *index_tuples, uint64_t num_windows, const double *windows, uint32_t options,
    double *d_result) {
    if (stat_id < 0 || stat_id > STAT_F4) return TSKB_ERR_BAD_PARAM_VALUE;
    return sample_count_stat(self, stat_id, num_sample_sets, sample_set_sizes, d_sample_sets,
        true, nullptr, num_index_tuples, index_tuples, num_sample_sets, 0, nullptr, num_windows,
        windows, options, d_result, true);
}

int tskb_treeseq_trees_at(const tskb_treeseq_t *self, uint64_t num_positions,
    const double *positions, const int32_t *tracked, uint64_t num_tracked, int32_t *out_parent,
    int32_t *out_count) {
    if (self == nullptr || self->plan == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    const Plan &P = *self->plan;
    for (uint64_t q = 0; q < num_positions; q++) {
        if (!(positions[q] >= 0 && positions[q] < P.L)) return TSKB_ERR_BAD_PARAM_VALUE;
    }
    for (uint64_t j = 0; tracked != nullptr && j < num_tracked; j++) {
        if (tracked[j] < 0 || tracked[j] >= (int32_t) P.N) return TSKB_ERR_NODE_OUT_OF_BOUNDS;
    }
    return guarded([&]() -> int {
        return run_trees_at(&P, num_positions, positions, tracked, num_tracked, out_parent,
            out_count);
    });
}

int tskb_treeseq_divergence_matrix(const tskb_treeseq_t *self, uint64_t num_sample_sets,
    const uint64_t *sample_set_sizes, const int32_t *sample_sets, uint64_t num_windows,
    const double *windows, uint32_t options, double *result) {
    if (self == nullptr || self->plan == nullptr || result == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    return guarded([&]() -> int {
        return run_divergence_matrix(self->plan, num_sample_sets, sample_set_sizes, sample_sets,
            num_windows, windows, options, result);
    });
}

int tskb_treeseq_genotype_matrix(const tskb_treeseq_t *self, const int32_t *samples,
    uint64_t num_samples, uint32_t options, int8_t *genotypes) {
    if (self == nullptr || self->plan == nullptr || genotypes == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    return guarded([&]() -> int {
        return run_genotype_matrix(self->plan, samples, num_samples, options, genotypes, 0, self->plan->S);
    });
}

int tskb_treeseq_decode_sites(const tskb_treeseq_t *self, uint64_t first_site, uint64_t num_sites,
    const int32_t *samples, uint64_t num_samples, uint32_t options, int8_t *genotypes) {
    if (self == nullptr || self->plan == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    if (num_sites > 0 && genotypes == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    return guarded([&]() -> int {
        return run_genotype_matrix(self->plan, samples, num_samples, options, genotypes, first_site, num_sites);
    });
}

int tskb_treeseq_get_stats(const tskb_treeseq_t *self, tskb_stats_t *out) {
    if (self == nullptr || self->plan == nullptr || out == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    std::lock_guard<std::mutex> lock(self->plan->mu);
    *out = self->plan->stats;
    return 0;
}

int64_t tskb_treeseq_debug_array(const tskb_treeseq_t *self, const char *name, void *out,
    uint64_t max_bytes) {
    if (self == nullptr || self->plan == nullptr || name == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    const Plan &P = *self->plan;
    std::lock_guard<std::mutex> lock(P.mu);
    const void *src = nullptr;
    size_t n = 0, esize = 0;
    bool host = false;
    std::string s(name);
#define ARR(NAME, A) if (s == NAME) { src = P.A.p; n = P.A.n; esize = sizeof(*P.A.p); }
    ARR("ev_pos", ev_pos) ARR("ev_child", ev_child) ARR("ev_sign", ev_sign) ARR("voff", voff)
    ARR("q_off", q_off) ARR("refs", refs) ARR("q_bp0", q_bp0) ARR("q_bp1", q_bp1) ARR("q_bl", q_bl) ARR("bp_pos", bp_pos)
    ARR("tile_dep", tile_dep)
    ARR("level", level) ARR("rank_node", rank_node) ARR("mut_src", mut_src)
    ARR("mut_allele", mut_allele) ARR("mut_alt", mut_alt)
#undef ARR
    if (s == "trace" && P.stats_trace != nullptr) {
        src = P.stats_trace; n = (size_t) P.ntiles * 4; esize = sizeof(unsigned long long);
    }
    if (s == "level_begin") {
        src = P.level_begin.data(); n = P.level_begin.size(); esize = sizeof(uint32_t); host = true;
    }
    if (esize == 0) return TSKB_ERR_BAD_PARAM_VALUE;
    size_t bytes = std::min<size_t>(n * esize, max_bytes);
    if (out != nullptr && bytes > 0) {
        if (host) {
            memcpy(out, src, bytes);
        } else {
            cudaSetDevice(P.device);
            if (cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) {
                cudaGetLastError();
                return TSKB_ERR_CUDA;
            }
        }
    }
    return (int64_t) n;
}

}  // extern "C"
