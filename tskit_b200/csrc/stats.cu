// stats.cu -- per-call kernels of the sample-count statistics
// (tsk_treeseq_sample_count_stat -> tsk_treeseq_general_stat, c/tskit/trees.c:2035-2220).
//
// Phases of one call (all on the plan's stream):
//   0 weights    sample sets -> 0/1 int32 weight rows per node          (trees.c:2195-2213)
//   1 sweep      ONE launch: state[u] over every piece = sum of the states it references; tiles of
//                a height wait on a completion counter for the heights below
//                                                                       (trees.c:1317-1327)
//                branch: the same launch turns every piece into window contributions
//                G = branch_length * f(state) over [x, xe)              (trees.c:1339-1350, 1484-1504)
//   2 summary    branch: only when the window bins do not fit shared memory (second pass)
//                site:   per site, allele states and sum of f           (trees.c:1525-1652)
//   3 finalize   branch: prefix over windows + span-normalise; site: window sums
//                                                                       (trees.c:1753-1762, 1920-1934)
//   5 d2h        result -> host
#include <chrono>
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "plan.cuh"

namespace tskb {
namespace {

constexpr int TB = 256;

// State of a node: one column per sample set (int32 counts) or per weight column (fp64 sums).
template <typename T, int K>
struct alignas(sizeof(T) * K >= 16 ? 16 : sizeof(T) * K) SVec {
    using scalar = T;
    static constexpr int N = K;
    T v[K];
    __host__ __device__ __forceinline__ SVec operator+(const SVec &o) const {
        SVec r;
#pragma unroll
        for (int k = 0; k < K; k++) r.v[k] = v[k] + o.v[k];
        return r;
    }
    __host__ __device__ __forceinline__ SVec operator-(const SVec &o) const {
        SVec r;
#pragma unroll
        for (int k = 0; k < K; k++) r.v[k] = v[k] - o.v[k];
        return r;
    }
};
template <int K> using IVec = SVec<int32_t, K>;
template <int K> using DVec = SVec<double, K>;

template <class V>
__host__ __device__ __forceinline__ V ivec_zero() {
    V r;
#pragma unroll
    for (int k = 0; k < V::N; k++) r.v[k] = 0;
    return r;
}

// statistics whose column reads only the (at most four) sets of its own index tuple: the lane's
// column then works on a 4-column excerpt of the state, fetched from shared memory
template <int STAT>
constexpr bool stat_reads_tuple_only() {
    return STAT == STAT_DIVERSITY || STAT == STAT_SEGSITES || STAT == STAT_Y1 || STAT == STAT_DIVERGENCE
           || STAT == STAT_Y2 || STAT == STAT_F2 || STAT == STAT_RELATEDNESS_NC || STAT == STAT_Y3
           || STAT == STAT_F3 || STAT == STAT_F4;
}

// one result column: the sample-set indexes of its tuple and their sizes
struct ColP {
    int32_t i, j, k, l;
    double ni, nj, nk, nl;
    double inv;  // 1 / denominator of the column (a constant of the sample set sizes)
};

struct SumP {
    int K;
    int M;
    int polarised;
    int skip_zero_bl;       // every summary value is finite: pieces with zero branch length add nothing
    double n[8];            // sample set sizes
    const ColP *cols;       // device [M]
    const double *table;    // device [rows * M] (STAT_TABULATED)
    uint32_t table_rows;
    int timing_hack;        // experiments (TSKB_SUM_HACK): 1 = leave out the run-head reductions (WRONG results)
};

// x[i] without dynamic register indexing, in the state's own type
template <class V>
__device__ __forceinline__ typename V::scalar pick_t(const V &s, int i) {
    typename V::scalar r = s.v[0];
#pragma unroll
    for (int k = 1; k < V::N; k++) r = (i == k) ? s.v[k] : r;
    return r;
}

// totals - s, evaluated lazily per picked column (see F_branch)
template <class V>
struct Compl {
    const V &s;
    const V &t;
    static constexpr int N = V::N;
    using scalar = typename V::scalar;
    using inner = V;
};
template <class V> struct is_compl { static constexpr bool value = false; };
template <class V> struct is_compl<Compl<V>> { static constexpr bool value = true; };

// x[i] without dynamic register indexing
template <class V>
__device__ __forceinline__ double pick(const V &s, int i) {
    if constexpr (is_compl<V>::value) {
        return (double) pick_t<typename V::inner>(s.t, i) - (double) pick_t<typename V::inner>(s.s, i);
    } else {
        typename V::scalar r = s.v[0];
#pragma unroll
        for (int k = 1; k < V::N; k++) r = (i == k) ? s.v[k] : r;
        return (double) r;
    }
}

// Denominator of a column, in the reference's operation order.  The device multiplies by its
// reciprocal: 0 * inf is NaN and x * inf is +-inf exactly where the reference's 0/0 and x/0 are
// (sample sets of size 1 or 2, test_stats.c:1737-1741), and a zero numerator stays an exact 0.
inline double column_denominator(int stat, const ColP &c) {
    switch (stat) {
        case STAT_DIVERSITY: return c.ni * (c.ni - 1);
        case STAT_Y1: return c.ni * (c.ni - 1) * (c.ni - 2);
        case STAT_DIVERGENCE: return c.ni * (c.nj - (c.i == c.j));
        case STAT_Y2: return c.ni * c.nj * (c.nj - 1);
        case STAT_F2: return c.ni * (c.ni - 1) * c.nj * (c.nj - 1);
        case STAT_RELATEDNESS_NC: return c.ni * c.nj;
        case STAT_Y3: return c.ni * c.nj * c.nk;
        case STAT_F3: return c.ni * (c.ni - 1) * c.nj * c.nk;
        case STAT_F4: return c.ni * c.nj * c.nk * c.nl;
    }
    return 1.0;
}

// The summary functions, numerators in the reference's exact operation order
// (c/tskit/trees.c:3934-3948, 4221-4264, 4690-4773, 4899-4959, 5177-5291).
template <int STAT, class V>
__device__ __forceinline__ double f_eval(const SumP &P, const ColP &c, int m, const V &s) {
    if constexpr (STAT == STAT_DIVERSITY) {
        double n = c.ni, x = pick<V>(s, c.i);
        return x * (n - x) * c.inv;
    } else if constexpr (STAT == STAT_SEGSITES) {
        double n = c.ni, x = pick<V>(s, c.i);
        return (x > 0) * (1 - x / n);
    } else if constexpr (STAT == STAT_Y1) {
        double ni = c.ni, xi = pick<V>(s, c.i);
        double numer = xi * (ni - xi) * (ni - xi - 1);
        return numer * c.inv;
    } else if constexpr (STAT == STAT_DIVERGENCE) {
        return pick<V>(s, c.i) * (c.nj - pick<V>(s, c.j)) * c.inv;
    } else if constexpr (STAT == STAT_Y2) {
        double nj = c.nj;
        double xi = pick<V>(s, c.i), xj = pick<V>(s, c.j);
        return xi * (nj - xj) * (nj - xj - 1) * c.inv;
    } else if constexpr (STAT == STAT_F2) {
        double ni = c.ni, nj = c.nj;
        double xi = pick<V>(s, c.i), xj = pick<V>(s, c.j);
        double numer = xi * (xi - 1) * (nj - xj) * (nj - xj - 1) - xi * (ni - xi) * (nj - xj) * xj;
        return numer * c.inv;
    } else if constexpr (STAT == STAT_RELATEDNESS) {
        double sumx = 0;
#pragma unroll
        for (int k = 0; k < V::N; k++) {
            if (k < P.K) sumx += pick<V>(s, k) / P.n[k];
        }
        double meanx = sumx / (double) P.K;
        return (pick<V>(s, c.i) / c.ni - meanx) * (pick<V>(s, c.j) / c.nj - meanx);
    } else if constexpr (STAT == STAT_RELATEDNESS_NC) {
        return pick<V>(s, c.i) * pick<V>(s, c.j) * c.inv;
    } else if constexpr (STAT == STAT_Y3) {
        double numer = pick<V>(s, c.i) * (c.nj - pick<V>(s, c.j)) * (c.nk - pick<V>(s, c.k));
        return numer * c.inv;
    } else if constexpr (STAT == STAT_F3) {
        double ni = c.ni, nj = c.nj, nk = c.nk;
        double xi = pick<V>(s, c.i), xj = pick<V>(s, c.j), xk = pick<V>(s, c.k);
        double numer = xi * (xi - 1) * (nj - xj) * (nk - xk) - xi * (ni - xi) * (nj - xj) * xk;
        return numer * c.inv;
    } else if constexpr (STAT == STAT_F4) {
        double nj = c.nj, nk = c.nk, nl = c.nl;
        double xi = pick<V>(s, c.i), xj = pick<V>(s, c.j), xk = pick<V>(s, c.k),
               xl = pick<V>(s, c.l);
        double numer = xi * xk * (nj - xj) * (nl - xl) - xi * xl * (nj - xj) * (nk - xk);
        return numer * c.inv;
    } else if constexpr (STAT == STAT_TRAIT_COV) {
        // trees.c:3960-3974; c.ni = 2 (n - 1)(n - 1) in the reference's operation order
        const double x = pick<V>(s, c.i);
        return (x * x) / c.ni;
    } else if constexpr (STAT == STAT_TRAIT_CORR) {
        // trees.c:4031-4049: the last state column is the frequency of the samples below
        const double n = P.n[0], x = pick<V>(s, c.i), p = pick<V>(s, P.K - 1);
        if ((p > 0.0) && (p < 1.0)) return (x * x) / (2 * (p * (1 - p)) * n * (n - 1));
        return 0.0;
    } else if constexpr (STAT == STAT_REL_WEIGHTED) {
        // trees.c:4800-4820: c.ni, c.nj = total weights of the two columns
        const double pn = pick<V>(s, P.K - 1);
        return (pick<V>(s, c.i) - c.ni * pn) * (pick<V>(s, c.j) - c.nj * pn);
    } else if constexpr (STAT == STAT_TRAIT_LM) {
        // trees.c:4106-4151: columns [0, M) traits, [M, K - 1) covariates, K - 1 the number of samples
        // below; P.table = V [M x table_rows] (traits^T covariates), table_rows = number of covariates
        const double num_samples = P.n[0], mm = pick<V>(s, P.K - 1);
        if ((mm > 0.0) && (mm < num_samples)) {
            double a = pick<V>(s, c.i), denom = mm;
            for (uint32_t j = 0; j < P.table_rows; j++) {
                const double z = pick<V>(s, P.M + (int) j);
                a -= z * __ldg(P.table + (size_t) c.i * P.table_rows + j);
                denom -= z * z;
            }
            if (denom < 1e-8) return 0.0;
            return (a * a) / (2 * denom * denom);
        }
        return 0.0;
    } else if constexpr (STAT == STAT_REL_SIDE) {
        // trees.c:4729-4753 with the mean over all table_rows sample sets carried as the last state
        // column (sum over the samples below of sum_k [sample in set k] / n_k)
        const double meanx = pick<V>(s, P.K - 1) / (double) P.table_rows;
        return (pick<V>(s, c.i) / c.ni - meanx) * (pick<V>(s, c.j) / c.nj - meanx);
    } else if constexpr (STAT == STAT_REL_WEIGHTED_NC) {
        return pick<V>(s, c.i) * pick<V>(s, c.j);  // trees.c:4822-4838
    } else {  // STAT_TABULATED
        uint32_t cnt = (uint32_t) (long long) pick<V>(s, 0);
        if (cnt >= P.table_rows) cnt = P.table_rows - 1;
        return __ldg(P.table + (size_t) cnt * P.M + m);
    }
}

// branch mode: f(x) + f(total - x) unless polarised (trees.c:1944-1972)
template <int STAT, class V>
__device__ __forceinline__ double F_branch(const SumP &P, const ColP &c, int m, const V &s,
    const V &totals) {
    double r = f_eval<STAT, V>(P, c, m, s);
    if (!P.polarised) r += f_eval<STAT, V>(P, c, m, totals - s);
    return r;
}

// ---------------------------------------------------------------- phase 0

// sample sets -> 0/1 weight columns of the samples' INIT slots (trees.c:2195-2213), and the
// argument checks of tsk_treeseq_check_sample_sets (trees.c:2114-2149) and of the duplicate scan
// (trees.c:2201-2214) on the way: verr = smallest (position << 1 | kind) of an element that is out
// of bounds (kind 0) or not a sample (kind 1) -- the reference reports the first one in order --
// and dup = some sample listed twice in one set.
template <class V>
__global__ void k_set_weights(const int32_t *sets, const uint32_t *set_off, uint32_t K,
    uint32_t total, const int32_t *sample_index, int32_t N, V *init, unsigned long long *verr,
    int *dup) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= total) return;
    uint32_t k = upper_bound_dev(set_off, K + 1, j) - 1;
    const int32_t u = sets[j];
    if (u < 0 || u >= N) {
        atomicMin(verr, (unsigned long long) j << 1);
        return;
    }
    const int32_t si = sample_index[u];
    if (si < 0) {
        atomicMin(verr, ((unsigned long long) j << 1) | 1ull);
        return;
    }
    // a sample may be in several sets (trees.c:2201-2214): distinct columns
    if (atomicExch(&init[si].v[k], 1) != 0) *dup = 1;
}

// weighted statistics: the samples' INIT slots hold their weight rows (trees.c:1406-1415 with
// general_stat's W)
template <class V>
__global__ void k_init_weights(const double *W, uint32_t n, uint32_t K, V *init) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    V r = ivec_zero<V>();
#pragma unroll
    for (int k = 0; k < V::N; k++) {
        if ((uint32_t) k < K) r.v[k] = (typename V::scalar) W[(size_t) i * K + k];
    }
    init[i] = r;
}

// ---------------------------------------------------------------- branch-mode running sum
// The reference keeps a running sum  S = sum over nodes of branch_length[u] * F(state[u]),
// updates it at every edge diff, and adds S * (distance to the next breakpoint or window edge) to
// the current window (trees.c:1339-1350, 1424-1507).  Here every piece [bp0, bp1) of a node adds
// +G at breakpoint bp0 and -G at breakpoint bp1, G = branch_length * F(state), to the array D of
// per-breakpoint deltas of S (native fp64 reductions in L2, spread over millions of addresses:
// no shared-memory compare-and-swap loops, no window lookup per piece).  S is the prefix sum of D
// and the windows integrate S (k_window_integrate).  NaN/inf summary values behave as in the
// reference: S is NaN from the breakpoint that first meets one.

struct DeltaOut {
    double *D;        // [M][Tp1]
    uint32_t Tp1;     // breakpoints + 1 (slot T = end of the range)
    const ColP *cols; // [M]
};

// the NP pieces one thread holds (lane-adjacent in the processing order) -> D, columns outermost
// (one column in registers at a time).  Where a piece ends at the breakpoint the next lane's
// piece starts at -- consecutive pieces of one node, mostly -- the two reductions to that
// address are merged into one of G_next - G (the reference's "subtract the old summary, add the
// new one" of one node at one breakpoint).  Must be called by all 32 lanes of a warp.
template <int STAT, class V, int NP>
__device__ __forceinline__ void pieces_to_deltas(const SumP &sp, const V &totals,
    const V (&st)[NP], const double (&bl)[NP], const uint32_t (&bp0)[NP],
    const uint32_t (&bp1)[NP], const DeltaOut &out, uint32_t m0, uint32_t m1, const ColP &first_col) {
    const uint32_t lane = threadIdx.x & 31u;
    bool valid[NP], live[NP], merge_next[NP], merged_prev[NP];
#pragma unroll
    for (int q = 0; q < NP; q++) {
        valid[q] = bp1[q] != NO_PIECE;  // not padding of the processing order
        // when every summary value is finite, roots and detached nodes add nothing: 0 * f (with
        // NaN/inf summaries 0 * f is NaN in the reference too, trees.c:1339-1350: kept then)
        live[q] = valid[q] && !(sp.skip_zero_bl && bl[q] == 0.0);
        const uint32_t next_bp0 = __shfl_down_sync(0xffffffffu, bp0[q], 1);
        const uint32_t prev_bp1 = __shfl_up_sync(0xffffffffu, bp1[q], 1);
        const bool next_valid = __shfl_down_sync(0xffffffffu, (int) valid[q], 1) != 0;
        merge_next[q] = valid[q] && lane < 31u && next_valid && next_bp0 == bp1[q];
        merged_prev[q] = valid[q] && lane > 0u && prev_bp1 == bp0[q];  // prev_bp1 of padding never matches
    }
    for (uint32_t m = m0; m < m1; m++) {
        const ColP col = m == m0 ? first_col : out.cols[m];  // the first column stays in registers
        double *Dm = out.D + (size_t) (m - m0) * out.Tp1;
#pragma unroll
        for (int q = 0; q < NP; q++) {
            double G = 0.0;
            if (live[q]) G = bl[q] * F_branch<STAT, V>(sp, col, m, st[q], totals);
            const double G_next = __shfl_down_sync(0xffffffffu, G, 1);
            if (!valid[q]) continue;
            if (!merged_prev[q] && G != 0.0 && !sp.timing_hack) atomicAdd(Dm + bp0[q], G);
            const double v = merge_next[q] ? G_next - G : -G;
            if (v != 0.0) atomicAdd(Dm + bp1[q], v);
        }
    }
}

// ---------------------------------------------------------------- phase 1 (+ 2, branch mode)
// state[u] over every piece (what update_state, trees.c:1317-1327, maintains incrementally):
// a piece's state is the sum of the states it references -- its children in the tree right of
// its breakpoint, plus its own INIT slot (the sample weight, trees.c:1406-1415).
// Pieces are processed by height; one cooperative launch of co-resident persistent CTAs covers
// every height.  CTA b takes tiles b, b + G, ...; a tile loads its references, then waits on
// the completion counter until every tile of the lower heights has published its states
// (red.release / ld.acquire at gpu scope), gathers, sums, stores (coalesced: the state slot of a
// piece is its processing position) and publishes.  Every CTA processes its tiles in increasing
// order and a tile only waits on lower-numbered tiles, so the lowest unfinished tile can always
// run: no deadlock.  Integer sums: exact in any order.
// FUSE: after publishing, the tile's pieces are turned into window contributions straight from
// registers (shared-memory bins, flushed once per CTA) -- the branch summary costs no second
// pass over the states.

template <class V>
__device__ __forceinline__ V state_load(const V *p) {
    // states are written by other CTAs of the same launch: read through L2, never L1
    union { V val; int4 q[sizeof(V) >= 16 ? sizeof(V) / 16 : 1]; int2 d; int w; } u;
    if constexpr (sizeof(V) >= 16) {
        const int4 *q = reinterpret_cast<const int4 *>(p);
#pragma unroll
        for (int c = 0; c < (int) (sizeof(V) / 16); c++) u.q[c] = __ldcg(q + c);
    } else if constexpr (sizeof(V) == 8) {
        u.d = __ldcg(reinterpret_cast<const int2 *>(p));
    } else {
        u.w = __ldcg(reinterpret_cast<const int *>(p));
    }
    return u.val;
}

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(uint32_t *p, uint32_t v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// optional per-tile timeline (TSKB_TRACE=1): 4 timestamps per tile
#define TRACE(slot) do { if (trace != nullptr && threadIdx.x == 0) trace[(size_t) tile * 4 + (slot)] = gtime(); } while (0)

// threads per sweep CTA (a tile is always PROP_TILE pieces): 512 for small states (measured on C2,
// ms per step: 128: 0.696, 256: 0.667, 512: 0.650, 1024: 0.871), 256 where a state is 16 bytes or more
template <class V> constexpr int prop_tb() { return sizeof(V) <= 8 ? 512 : 256; }
// references fetched before the wait: 4 for small states (measured on C2: 2: 0.729, 3: 0.692, 4: 0.663,
// 5: 0.685, 6: 0.725 ms per step), 3 where a state is 16 bytes or more (registers); the rare piece
// with more (multifurcations) finishes in a loop
template <class V> constexpr int prop_pre() { return sizeof(V) <= 8 ? 4 : 3; }  // references fetched before the wait; more are rare (multifurcations)
constexpr uint32_t SPIN_LIMIT = 1u << 22;  // ~seconds; a legitimate wait is < the kernel's own run time

struct SweepArgs {
    uint32_t ntiles;
    const uint32_t *tile_dep, *q_off, *refs;
    uint32_t *counters;
    int *error_flag;
    unsigned long long *trace;
};

template <class V>
__global__ void __launch_bounds__(prop_tb<V>()) k_sweep(SweepArgs a, V *pval) {
    constexpr int PROP_PRE = prop_pre<V>();
    constexpr int PROP_TB = prop_tb<V>(), PROP_IPT = PROP_TILE / PROP_TB;
    unsigned long long *trace = a.trace;
    for (uint32_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        TRACE(0);
        const uint32_t dep = __ldg(a.tile_dep + tile);
        const uint32_t j0 = tile * PROP_TILE + threadIdx.x;
        uint32_t o0[PROP_IPT], o1[PROP_IPT], rf[PROP_IPT][PROP_PRE];
        // everything that does not depend on other tiles is fetched before the wait
#pragma unroll
        for (int q = 0; q < PROP_IPT; q++) {
            o0[q] = __ldg(a.q_off + j0 + q * PROP_TB);
            o1[q] = __ldg(a.q_off + j0 + q * PROP_TB + 1);
        }
#pragma unroll
        for (int q = 0; q < PROP_IPT; q++) {
#pragma unroll
            for (int i = 0; i < PROP_PRE; i++) {
                rf[q][i] = o0[q] + i < o1[q] ? __ldg(a.refs + o0[q] + i) : NO_PIECE;
            }
        }
        if (dep > 0) {
            if (threadIdx.x == 0) {
                uint32_t spins = 0;
                while (ld_acquire(a.counters + 1) < dep) {
                    if (++spins > SPIN_LIMIT || ((spins & 1023u) == 0 && *(volatile int *) a.error_flag)) {
                        *a.error_flag = 1;
                        break;
                    }
                }
            }
            __syncthreads();
        }
        TRACE(1);
        V g[PROP_IPT][PROP_PRE];
#pragma unroll
        for (int q = 0; q < PROP_IPT; q++) {
#pragma unroll
            for (int i = 0; i < PROP_PRE; i++) {
                g[q][i] = ivec_zero<V>();
                if (rf[q][i] != NO_PIECE) g[q][i] = state_load<V>(pval + rf[q][i]);
            }
        }
        TRACE(2);
#pragma unroll
        for (int q = 0; q < PROP_IPT; q++) {
            V sum = g[q][0];
#pragma unroll
            for (int i = 1; i < PROP_PRE; i++) sum = sum + g[q][i];
            for (uint32_t o = o0[q] + PROP_PRE; o < o1[q]; o++) {
                sum = sum + state_load<V>(pval + __ldg(a.refs + o));
            }
            pval[j0 + q * PROP_TB] = sum;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            red_release_add(a.counters + 1, 1u);
            if (trace != nullptr) trace[(size_t) tile * 4 + 3] = gtime();
        }
    }
}

// ---------------------------------------------------------------- phase 2, branch mode
// The states are streamed once (fully coalesced) per chunk of result columns; a chunk is as many
// columns as fit the delta budget (all of them, normally).  Kept separate from the sweep: both are
// bound by the rate at which an SM issues requests to L2 (measured: fusing them saved nothing), and
// apart they keep their own register budgets.

#ifndef TSKB_SUM_IPT
#define TSKB_SUM_IPT 1
#endif
#ifndef TSKB_SUM_TB
#define TSKB_SUM_TB 512
#endif
constexpr int SUM_IPT = TSKB_SUM_IPT;    // pieces per thread and pipeline stage
constexpr int SUM_TB = TSKB_SUM_TB;      // threads per summary CTA
constexpr int SUM_TILE = SUM_TB * SUM_IPT;

template <class V>
struct PieceRegs {
    V st[SUM_IPT];
    double bl[SUM_IPT];
    uint32_t bp0[SUM_IPT], bp1[SUM_IPT];
    __device__ __forceinline__ void load(uint32_t tile, uint32_t npp, const uint32_t *__restrict__ q_bp0,
        const uint32_t *__restrict__ q_bp1, const double *__restrict__ q_bl,
        const V *__restrict__ pval) {
#pragma unroll
        for (int q = 0; q < SUM_IPT; q++) {
            const uint32_t j = tile * SUM_TILE + q * SUM_TB + threadIdx.x;
            bp1[q] = NO_PIECE;  // past the end: padding
            bp0[q] = 0;
            bl[q] = 0.0;
            st[q] = ivec_zero<V>();
            // the processing order is padded to whole PROP_TILEs: no bounds check when tiles divide them
            if (PROP_TILE % SUM_TILE == 0 || j < npp) {
                st[q] = pval[j]; bl[q] = q_bl[j]; bp0[q] = q_bp0[j]; bp1[q] = q_bp1[j];
            }
        }
    }
};

// Few warps with many loads in flight each beat many warps here (measured): the kernel is bound
// by the rate at which an SM can issue requests to L2, and the reductions of a warp are 32
// separate requests.
template <int STAT, class V>
__global__ void __launch_bounds__(SUM_TB) k_branch_summary(uint32_t npp,
    const uint32_t *__restrict__ q_bp0, const uint32_t *__restrict__ q_bp1,
    const double *__restrict__ q_bl, const V *__restrict__ pval, SumP sp, V totals,
    DeltaOut out, uint32_t m0, uint32_t m1) {
    const uint32_t ntiles = (npp + SUM_TILE - 1) / SUM_TILE;
    // software pipeline: the next tile's loads are in flight while this one is evaluated
    PieceRegs<V> cur, nxt;
    const ColP first_col = out.cols[m0];
    uint32_t tile = blockIdx.x;
    if (tile < ntiles) cur.load(tile, npp, q_bp0, q_bp1, q_bl, pval);
    for (; tile < ntiles; tile += gridDim.x) {
        const uint32_t tn = tile + gridDim.x;
        if (tn < ntiles) nxt.load(tn, npp, q_bp0, q_bp1, q_bl, pval);
        pieces_to_deltas<STAT, V, SUM_IPT>(sp, totals, cur.st, cur.bl, cur.bp0, cur.bp1, out, m0, m1, first_col);
        cur = nxt;
    }
}

// ---- thread-contiguous variant (default): a thread holds SUMC_IPT consecutive pieces of the processing
// order, fetched with 16-byte loads, and merges the reductions of abutting pieces in its own registers;
// only the boundary between two lanes goes through the shuffle network (4 SHFL per 4 pieces instead
// of 6 per piece: the shuffles share the MIO queue with the reductions).  Same reductions, same
// addresses and values as pieces_to_deltas.
#ifndef TSKB_SUMC_TB
#define TSKB_SUMC_TB 256
#endif
constexpr int SUMC_TB = TSKB_SUMC_TB;
constexpr int SUMC_IPT = 4;

template <class V>
struct Quad {
    V st[SUMC_IPT];
    double bl[SUMC_IPT];
    uint32_t bp0[SUMC_IPT], bp1[SUMC_IPT];
    __device__ __forceinline__ void load(uint32_t qd, const uint32_t *__restrict__ q_bp0,
        const uint32_t *__restrict__ q_bp1, const double *__restrict__ q_bl, const V *__restrict__ pval) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(q_bp0) + qd);
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(q_bp1) + qd);
        bp0[0] = a.x; bp0[1] = a.y; bp0[2] = a.z; bp0[3] = a.w;
        bp1[0] = b.x; bp1[1] = b.y; bp1[2] = b.z; bp1[3] = b.w;
        const double2 l0 = __ldg(reinterpret_cast<const double2 *>(q_bl) + 2 * (size_t) qd);
        const double2 l1 = __ldg(reinterpret_cast<const double2 *>(q_bl) + 2 * (size_t) qd + 1);
        bl[0] = l0.x; bl[1] = l0.y; bl[2] = l1.x; bl[3] = l1.y;
        constexpr int NQ = (int) (sizeof(V) * SUMC_IPT / 16);
        union { V s[SUMC_IPT]; int4 q[NQ]; } u;
        const int4 *src = reinterpret_cast<const int4 *>(pval + (size_t) qd * SUMC_IPT);
#pragma unroll
        for (int i = 0; i < NQ; i++) u.q[i] = src[i];
#pragma unroll
        for (int i = 0; i < SUMC_IPT; i++) st[i] = u.s[i];
    }
};

template <int STAT, class V>
__global__ void __launch_bounds__(SUMC_TB) k_branch_summary_c4(uint32_t npp,
    const uint32_t *__restrict__ q_bp0, const uint32_t *__restrict__ q_bp1,
    const double *__restrict__ q_bl, const V *__restrict__ pval, SumP sp, V totals,
    DeltaOut out, uint32_t m0, uint32_t m1) {
    // npp is a multiple of PROP_TILE = 1024: whole warps of quads, no bounds checks
    const uint32_t nquads = npp / SUMC_IPT;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t stride = gridDim.x * blockDim.x;
    const ColP first_col = out.cols[m0];
    uint32_t qd = blockIdx.x * blockDim.x + threadIdx.x;
    Quad<V> cur, nxt;
    if (qd < nquads) cur.load(qd, q_bp0, q_bp1, q_bl, pval);
    for (; qd < nquads; qd += stride) {
        const uint32_t qn = qd + stride;
        if (qn < nquads) nxt.load(qn, q_bp0, q_bp1, q_bl, pval);
        bool valid[SUMC_IPT], live[SUMC_IPT];
#pragma unroll
        for (int q = 0; q < SUMC_IPT; q++) {
            valid[q] = cur.bp1[q] != NO_PIECE;
            live[q] = valid[q] && !(sp.skip_zero_bl && cur.bl[q] == 0.0);
        }
        // the neighbouring lanes' abutting pieces; padding has bp0 = 0 and bp1 = NO_PIECE, which never
        // match a real piece's bp1 >= 1 and bp0 <= T
        uint32_t prev_bp1 = __shfl_up_sync(0xffffffffu, cur.bp1[SUMC_IPT - 1], 1);
        uint32_t next_bp0 = __shfl_down_sync(0xffffffffu, cur.bp0[0], 1);
        if (lane == 0u) prev_bp1 = NO_PIECE;
        if (lane == 31u) next_bp0 = NO_PIECE;
        for (uint32_t m = m0; m < m1; m++) {
            const ColP col = m == m0 ? first_col : out.cols[m];
            double *Dm = out.D + (size_t) (m - m0) * out.Tp1;
            double G[SUMC_IPT];
#pragma unroll
            for (int q = 0; q < SUMC_IPT; q++) {
                G[q] = live[q] ? cur.bl[q] * F_branch<STAT, V>(sp, col, m, cur.st[q], totals) : 0.0;
            }
            const double G_left = __shfl_up_sync(0xffffffffu, G[SUMC_IPT - 1], 1);
#pragma unroll
            for (int q = 0; q < SUMC_IPT; q++) {
                if (!valid[q]) continue;
                const bool merged_prev = q == 0 ? prev_bp1 == cur.bp0[0] : cur.bp1[q > 0 ? q - 1 : 0] == cur.bp0[q];
                const double before = q == 0 ? G_left : G[q > 0 ? q - 1 : 0];
                const double v = merged_prev ? G[q] - before : G[q];
                if (v != 0.0) atomicAdd(Dm + cur.bp0[q], v);
                const bool merged_next = q == SUMC_IPT - 1 ? next_bp0 == cur.bp1[q]
                                                           : cur.bp0[q < SUMC_IPT - 1 ? q + 1 : q] == cur.bp1[q];
                if (!merged_next && G[q] != 0.0) atomicAdd(Dm + cur.bp1[q], -G[q]);
            }
        }
        cur = nxt;
    }
}

// ---- window bins in shared memory (few windows, finite summaries): no per-breakpoint deltas at all
// A piece [x0, x1) with G = branch_length x F(state) adds G x |[x0, x1) meet window| to every window.
// Written as two EVENTS -- (x0, +G) and (x1, -G) -- an event (x, v) in window w adds
//     C[w] += v,   P[w] -= v (x - left edge of w)
// and window w's integral is  P[w] + span(w) x (C[0] + ... + C[w])  (k_bins_finalize): the prefix of C is
// the running sum S at the window's right edge, P takes back what lies left of each event.  Abutting
// pieces of neighbouring lanes share an event of G_next - G.  The bins of a CTA live in shared memory
// (compare-and-swap additions: few and mostly uncontended), and are added to the global bins once per
// CTA: the scattered fp64 reductions in L2 of the delta formulation, which bound it at ~100 G/s,
// are gone, and so are the scan over the breakpoints and the window integration.
struct BinArgs {
    const double *windows;  // [W + 1] (device)
    uint32_t W;
    double w0, inv_width;   // uniform windows: index guess (x - w0) * inv_width, then corrected
    int uniform;
    double *gP, *gC;        // [ncols][W] global bins
};

__device__ __forceinline__ uint32_t bin_of(const double *win, const BinArgs &b, double x) {
    uint32_t w;
    if (b.uniform) {
        const double g = (x - b.w0) * b.inv_width;
        w = g <= 0.0 ? 0u : (g >= (double) b.W ? b.W : (uint32_t) g);
        while (w > 0 && x < win[w]) w--;
        while (w < b.W && x >= win[w + 1]) w++;
    } else {
        w = upper_bound_dev(win, b.W + 1, x);
        w = w > 0 ? w - 1 : 0;
    }
    return w;  // == W: at or beyond the last edge (no window)
}

__device__ __forceinline__ void bin_event(double *sP, double *sC, const double *win, const BinArgs &b,
    double x, double v) {
    const uint32_t w = bin_of(win, b, x);
    if (w < b.W) {
        atomicAdd(sC + w, v);
        atomicAdd(sP + w, -(v * (x - win[w])));
    }
}

template <class V>
struct PieceX {
    V st;
    double bl, x0, x1;
    __device__ __forceinline__ void load(uint32_t j, const double *__restrict__ q_x0,
        const double *__restrict__ q_x1, const double *__restrict__ q_bl, const V *__restrict__ pval) {
        st = pval[j]; bl = q_bl[j]; x0 = q_x0[j]; x1 = q_x1[j];
    }
};

template <int STAT, class V>
__global__ void __launch_bounds__(SUM_TB) k_branch_summary_bins(uint32_t npp,
    const double *__restrict__ q_x0, const double *__restrict__ q_x1, const double *__restrict__ q_bl,
    const V *__restrict__ pval, SumP sp, V totals, const ColP *cols, uint32_t m0, uint32_t ncols, BinArgs b) {
    extern __shared__ double smem_bins[];
    double *win = smem_bins;                       // [W + 1]
    double *sP = win + b.W + 1;                    // [ncols][W]
    double *sC = sP + (size_t) ncols * b.W;        // [ncols][W]
    for (uint32_t i = threadIdx.x; i <= b.W; i += blockDim.x) win[i] = b.windows[i];
    for (uint32_t i = threadIdx.x; i < 2 * ncols * b.W; i += blockDim.x) sP[i] = 0.0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t ntiles = npp / SUM_TB;  // npp is a multiple of PROP_TILE, itself a multiple of SUM_TB
    const ColP first_col = cols[m0];
    PieceX<V> cur, nxt;
    uint32_t tile = blockIdx.x;
    if (tile < ntiles) cur.load(tile * SUM_TB + threadIdx.x, q_x0, q_x1, q_bl, pval);
    for (; tile < ntiles; tile += gridDim.x) {
        const uint32_t tn = tile + gridDim.x;
        if (tn < ntiles) nxt.load(tn * SUM_TB + threadIdx.x, q_x0, q_x1, q_bl, pval);
        // padding has x1 < 0; pieces without a branch above them add nothing (finite summaries only here)
        const bool valid = cur.x1 >= 0.0;
        const bool live = valid && cur.bl != 0.0;
        const double next_x0 = __shfl_down_sync(0xffffffffu, cur.x0, 1);
        const double prev_x1 = __shfl_up_sync(0xffffffffu, cur.x1, 1);
        const bool merged_next = valid && lane < 31u && next_x0 == cur.x1;
        const bool merged_prev = valid && lane > 0u && prev_x1 == cur.x0;
        for (uint32_t c = 0; c < ncols; c++) {
            const uint32_t m = m0 + c;
            const ColP col = c == 0 ? first_col : cols[m];
            double G = 0.0;
            if (live) G = cur.bl * F_branch<STAT, V>(sp, col, m, cur.st, totals);
            const double G_prev = __shfl_up_sync(0xffffffffu, G, 1);
            if (!valid) continue;
            double *cP = sP + (size_t) c * b.W, *cC = sC + (size_t) c * b.W;
            const double v = merged_prev ? G - G_prev : G;
            if (v != 0.0) bin_event(cP, cC, win, b, cur.x0, v);
            if (!merged_next && G != 0.0) bin_event(cP, cC, win, b, cur.x1, -G);
        }
        cur = nxt;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < ncols * b.W; i += blockDim.x) {
        const double p = sP[i], q = sC[i];
        if (p != 0.0) atomicAdd(b.gP + i, p);
        if (q != 0.0) atomicAdd(b.gC + i, q);
    }
}

__global__ void k_piece_positions(uint32_t npp, const uint32_t *__restrict__ q_bp0,
    const uint32_t *__restrict__ q_bp1, const double *__restrict__ bp_pos, double *x0, double *x1) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npp) return;
    const uint32_t b1 = q_bp1[j];
    x0[j] = b1 == NO_PIECE ? 0.0 : bp_pos[q_bp0[j]];
    x1[j] = b1 == NO_PIECE ? -1.0 : bp_pos[b1];  // bp_pos[T] = end of the range
}

// window w of column c: P + span x (inclusive prefix of C); one warp per column
__global__ void k_bins_finalize(const double *__restrict__ gP, const double *__restrict__ gC,
    const double *__restrict__ windows, uint32_t W, uint32_t ncols, uint32_t m0, uint32_t M, int span_normalise,
    double *result) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= ncols) return;
    double run = 0.0;
    for (uint32_t base = 0; base < W; base += 32) {
        const uint32_t w = base + lane;
        double v = w < W ? gC[(size_t) c * W + w] : 0.0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double u = __shfl_up_sync(0xffffffffu, v, d);
            if ((int) lane >= d) v += u;
        }
        const double S = run + v;
        run += __shfl_sync(0xffffffffu, v, 31);
        if (w < W) {
            const double span = windows[w + 1] - windows[w];
            double r = gP[(size_t) c * W + w] + span * S;
            if (span_normalise) r /= span;
            result[(size_t) w * M + m0 + c] = r;
        }
    }
}

// ---- window runs in registers (few windows, finite summaries; the default for up to 5 columns)
// The window integral of a branch statistic is  sum over pieces of  G x |[x0, x1) meet window|,
// G = branch_length x F(state).  A thread walks RUN_IPT CONSECUTIVE pieces of the processing order
// (consecutive pieces are mostly successive pieces of one node: they follow each other along the genome)
// and keeps the sum for the window it is in IN A REGISTER; it leaves as one fp64 reduction when the
// thread moves to another window -- 0.3-0.5 reductions per piece instead of one per piece into the
// per-breakpoint deltas, and no scan over the breakpoints or window integration afterwards.  A piece
// that crosses window edges adds its parts to the first and the last window it meets and +-G to a
// difference array over the windows it covers entirely (gC).  The bins are replicated (one copy per
// group of warps: reductions to one address serialise in L2) and summed by k_runs_finalize.
#ifndef TSKB_RUN_TB
#define TSKB_RUN_TB 256
#endif
#ifndef TSKB_RUN_IPT
#define TSKB_RUN_IPT 8
#endif
#ifndef TSKB_RUN_MINB
#define TSKB_RUN_MINB 3
#endif
constexpr int RUN_TB = TSKB_RUN_TB;
constexpr int RUN_IPT = TSKB_RUN_IPT;   // consecutive pieces per thread (a multiple of 4: 16-byte loads)

struct RunArgs {
    const double *windows;  // [W + 1] (device)
    uint32_t W;
    double w0, inv_width;   // roughly uniform windows: index guess (x - w0) * inv_width, then corrected
    int uniform;
    double step, inv_step, wlast;  // exactly uniform windows: edge i = w0 + i * step (i < W), wlast (i = W)
    uint32_t wlo, Wl;       // the windows this engine's genome range can touch: [wlo, wlo + Wl); bins cover those
    double *bins;           // [copies][(2 Wl + 1) x ncols]: R[w - wlo][col] then C[w - wlo][col] (C has Wl + 1 rows)
    uint32_t ncols;         // columns per bin row
    uint32_t copy_mask;     // copies - 1 (a power of two)
    uint32_t chunk_mul;     // walk by start position: chunk of warp i = i * chunk_mul mod the number of chunks
};

// window w with win[w] <= x < win[w + 1] (END = false) or win[w] < x <= win[w + 1] (END = true);
// the caller guarantees win[0] <= x < win[W], resp. win[0] < x <= win[W]
template <bool END>
__device__ __forceinline__ uint32_t run_window_of(const double *win, const RunArgs &b, double x) {
    uint32_t lo = 0, hi = b.W - 1;
    if (b.uniform) {
        const double g = (x - b.w0) * b.inv_width;
        uint32_t w = g <= 0.0 ? 0u : (g >= (double) (b.W - 1) ? b.W - 1 : (uint32_t) g);
        if (END) {
            while (w > 0 && x <= win[w]) w--;
            while (w + 1 < b.W && x > win[w + 1]) w++;
        } else {
            while (w > 0 && x < win[w]) w--;
            while (w + 1 < b.W && x >= win[w + 1]) w++;
        }
        return w;
    }
    while (lo < hi) {  // largest w with win[w] <= x (resp. < x)
        const uint32_t mid = lo + ((hi - lo + 1) >> 1);
        const bool left = END ? win[mid] < x : win[mid] <= x;
        if (left) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <class V>
struct Run8 {
    V st[RUN_IPT];
    double bl[RUN_IPT], x0[RUN_IPT], x1[RUN_IPT];
    __device__ __forceinline__ void load(uint32_t g, const double *__restrict__ q_x0,
        const double *__restrict__ q_x1, const double *__restrict__ q_bl, const V *__restrict__ pval) {
        const size_t base = (size_t) g * RUN_IPT;
#pragma unroll
        for (int i = 0; i < RUN_IPT / 2; i++) {
            const double2 a = __ldg(reinterpret_cast<const double2 *>(q_bl + base) + i);
            const double2 b = __ldg(reinterpret_cast<const double2 *>(q_x0 + base) + i);
            const double2 c = __ldg(reinterpret_cast<const double2 *>(q_x1 + base) + i);
            bl[2 * i] = a.x; bl[2 * i + 1] = a.y;
            x0[2 * i] = b.x; x0[2 * i + 1] = b.y;
            x1[2 * i] = c.x; x1[2 * i + 1] = c.y;
        }
        constexpr int NQ = (int) (sizeof(V) * RUN_IPT / 16);
        union { V s[RUN_IPT]; int4 q[NQ]; } u;
        const int4 *src = reinterpret_cast<const int4 *>(pval + base);
#pragma unroll
        for (int i = 0; i < NQ; i++) u.q[i] = src[i];
#pragma unroll
        for (int i = 0; i < RUN_IPT; i++) st[i] = u.s[i];
    }
};

// Exactly uniform windows (np.linspace: edge i = w0 + i * step, the last one the stop value; the host has
// checked that every given edge equals this expression bit for bit -- no FMA contraction, -fmad=false):
// the window of a position is cell(x) = min(trunc((x - w0) / step), W - 1), a monotone function of x.
// Within a few ulp of an edge cell(x) may name the neighbouring window; the parts of a piece are
// measured against the nominal edges, so that they still add up to G x (e - a) exactly and the
// misplaced part is a few ulp of a base pair long.
template <int STAT, class V>
__global__ void __launch_bounds__(RUN_TB, TSKB_RUN_MINB) k_branch_summary_runs(uint32_t npp,
    const double *__restrict__ q_x0, const double *__restrict__ q_x1, const double *__restrict__ q_bl,
    const V *__restrict__ pval, SumP sp, V totals, const ColP *cols, uint32_t m, RunArgs b) {
    const double first_edge = b.w0, last_edge = b.wlast;
    const ColP col = cols[m];
    const uint32_t copy = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) & b.copy_mask;
    double *gR = b.bins + (size_t) copy * (2 * (size_t) b.Wl + 1) - b.wlo;   // indexed by the window itself
    double *gC = gR + b.Wl;
    // npp is a multiple of PROP_TILE = 1024: whole groups, no bounds checks inside a group
    const uint32_t ngroups = npp / RUN_IPT;
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t wc = 0;    // the window the thread's register sum belongs to
    double acc = 0.0;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
        Run8<V> cur;
        cur.load(g, q_x0, q_x1, q_bl, pval);
#pragma unroll
        for (int q = 0; q < RUN_IPT; q++) {
            // Straight-line arithmetic for all lanes up to the few predicated reductions: padding
            // (x1 < 0) and pieces without a branch above them (finite summaries only here) end up with
            // an empty clipped span or G = 0.
            const double G = cur.bl[q] * F_branch<STAT, V>(sp, col, m, cur.st[q], totals);
            // windows need not cover the genome (divergence_matrix): clip to [first_edge, last_edge]
            const double a = fmax(cur.x0[q], first_edge), e = fmin(cur.x1[q], last_edge);
            const bool live = a < e && G != 0.0;
            uint32_t w0, w1;
            double hi0, lo1;
            w0 = min((uint32_t) ((a - first_edge) * b.inv_step), b.W - 1);
            w1 = min((uint32_t) ((e - first_edge) * b.inv_step), b.W - 1);
            hi0 = w0 + 1 >= b.W ? last_edge : first_edge + (double) (w0 + 1) * b.step;
            lo1 = first_edge + (double) w1 * b.step;
            if (live) {
                if (w0 != wc) {  // the thread's register sum moves to the piece's first window
                    if (acc != 0.0) atomicAdd(gR + wc, acc);
                    acc = 0.0;
                    wc = w0;
                }
                if (w1 == w0) {
                    acc += G * (e - a);
                } else {  // a piece that ends in a later window closes the first one and opens the last one
                    atomicAdd(gR + w0, acc + G * (hi0 - a));
                    if (w1 > w0 + 1) {  // the windows in between are covered entirely
                        atomicAdd(gC + w0 + 1, G);
                        atomicAdd(gC + w1, -G);
                    }
                    wc = w1;
                    acc = G * (e - lo1);
                }
            }
        }
    }
    if (acc != 0.0) atomicAdd(gR + wc, acc);
}

// RC[i] = sum over the copies of bins[copy][i], i < block = (2 W + 1) x ncols; CTAs of 1024 threads:
// lanes are elements, the warps share out the copies
__global__ void __launch_bounds__(1024) k_runs_reduce(const double *__restrict__ bins, uint32_t block, uint32_t copies,
    double *RC) {
    __shared__ double part[32][33];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t i = blockIdx.x * 32 + lane;
    double s = 0.0;
    if (i < block) {
        for (uint32_t c = warp; c < copies; c += 32) s += bins[(size_t) c * block + i];
    }
    part[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && i < block) {
        double t = 0.0;
#pragma unroll 8
        for (int k = 0; k < 32; k++) t += part[k][lane];
        RC[i] = t;
    }
}

// window w of column c: R[w][c] + span(w) x (C[wlo][c] + ... + C[w][c]) for the windows [wlo, wlo + Wl) the bins
// cover, zero elsewhere; one CTA of 1024 threads per column: a thread sums a run of windows, the CTA scans
// the run totals, the thread walks its run again
__global__ void __launch_bounds__(1024) k_runs_finalize(const double *__restrict__ RC, const double *__restrict__ windows,
    uint32_t W, uint32_t wlo, uint32_t Wl, uint32_t ncols, uint32_t m0, uint32_t M, int span_normalise, double *result) {
    __shared__ double warp_tot[32];
    const uint32_t c = blockIdx.x, t = threadIdx.x, lane = t & 31u, wp = t >> 5;
    const double *R = RC + c, *C = RC + (size_t) Wl * ncols + c;
    for (uint32_t w = t; w < W; w += blockDim.x) {
        if (w < wlo || w >= wlo + Wl) result[(size_t) w * M + m0 + c] = 0.0;
    }
    const uint32_t per = (Wl + blockDim.x - 1) / blockDim.x;
    const uint32_t i0 = min(Wl, t * per), i1 = min(Wl, i0 + per);
    double sum = 0.0;
    for (uint32_t i = i0; i < i1; i++) sum += C[(size_t) i * ncols];
    double v = sum;  // inclusive scan over the threads
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, v, d);
        if ((int) lane >= d) v += u;
    }
    if (lane == 31u) warp_tot[wp] = v;
    __syncthreads();
    double before = 0.0;
    for (uint32_t q = 0; q < wp; q++) before += warp_tot[q];
    double S = before + v - sum;  // prefix of C before this thread's run
    for (uint32_t i = i0; i < i1; i++) {
        S += C[(size_t) i * ncols];
        const uint32_t w = wlo + i;
        const double span = windows[w + 1] - windows[w];
        double r = R[(size_t) i * ncols] + span * S;
        if (span_normalise) r /= span;
        result[(size_t) w * M + m0 + c] = r;
    }
}

// ---- many result columns: lanes are columns
// With M columns the kernel above issues M reductions per piece, each lane of an instruction to
// its own cache line.  Here a warp walks its pieces one after the other and lane m evaluates
// column m: the reductions of one piece go to M consecutive doubles (D is laid out
// [breakpoint][column] for this kernel), one or two L2 requests instead of M, and a piece that
// starts where the previous one ends is merged with it in registers.  Up to 32 columns per pass.
template <int STAT, class V>
__global__ void __launch_bounds__(TB) k_branch_summary_cols(uint32_t npp,
    const uint32_t *__restrict__ q_bp0, const uint32_t *__restrict__ q_bp1,
    const double *__restrict__ q_bl, const V *__restrict__ pval, SumP sp, V totals,
    DeltaOut out, uint32_t m0, uint32_t ncols) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t nchunks = (npp + 31) / 32;
    const bool mine = lane < ncols;
    const uint32_t m = m0 + (mine ? lane : 0);
    const ColP col = out.cols[m];
    double *Dl = out.D + lane;  // D[bp * ncols + lane]
    for (uint32_t chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; chunk < nchunks; chunk += warps) {
        const uint32_t j = chunk * 32 + lane;
        V st = ivec_zero<V>();
        double bl = 0.0;
        uint32_t bp0 = 0, bp1 = NO_PIECE;
        if (j < npp) {
            st = pval[j]; bl = q_bl[j]; bp0 = q_bp0[j]; bp1 = q_bp1[j];
        }
        double G_prev = 0.0;
        uint32_t end_prev = NO_PIECE;  // breakpoint the previous piece ends at; NO_PIECE: none pending
        for (int i = 0; i < 32; i++) {
            V s_i;
#pragma unroll
            for (int k = 0; k < V::N; k++) s_i.v[k] = __shfl_sync(0xffffffffu, st.v[k], i);
            const double bl_i = __shfl_sync(0xffffffffu, bl, i);
            const uint32_t b0 = __shfl_sync(0xffffffffu, bp0, i), b1 = __shfl_sync(0xffffffffu, bp1, i);
            double G = 0.0;
            const bool valid = b1 != NO_PIECE;  // warp-uniform
            if (valid && !(sp.skip_zero_bl && bl_i == 0.0)) G = bl_i * F_branch<STAT, V>(sp, col, m, s_i, totals);
            if (valid && end_prev == b0) {
                const double v = G - G_prev;
                if (mine && v != 0.0) atomicAdd(Dl + (size_t) b0 * ncols, v);
            } else {
                if (end_prev != NO_PIECE && mine && G_prev != 0.0) atomicAdd(Dl + (size_t) end_prev * ncols, -G_prev);
                if (valid && mine && G != 0.0) atomicAdd(Dl + (size_t) b0 * ncols, G);
            }
            G_prev = G;
            end_prev = valid ? b1 : NO_PIECE;
        }
        if (end_prev != NO_PIECE && mine && G_prev != 0.0) atomicAdd(Dl + (size_t) end_prev * ncols, -G_prev);
    }
}

// D[bp][col] -> Dt[col][bp]
__global__ void k_transpose_deltas(const double *__restrict__ D, uint32_t Tp1, uint32_t ncols, double *Dt) {
    __shared__ double tile[32][33];
    const uint32_t b0 = blockIdx.x * 32;
    for (uint32_t r = threadIdx.y; r < 32; r += blockDim.y) {
        const uint32_t bp = b0 + r;
        tile[r][threadIdx.x] = (bp < Tp1 && threadIdx.x < ncols) ? D[(size_t) bp * ncols + threadIdx.x] : 0.0;
    }
    __syncthreads();
    for (uint32_t cidx = threadIdx.y; cidx < ncols; cidx += blockDim.y) {
        const uint32_t bp = b0 + threadIdx.x;
        if (bp < Tp1) Dt[(size_t) cidx * Tp1 + bp] = tile[threadIdx.x][cidx];
    }
}

// ---- many result columns, pieces in the order of their start breakpoints (the summary order)
// With M columns the deltas are M x 8 bytes per breakpoint -- 1 GB for 28 columns on the C2 ARG, far
// past L2, so reductions scattered over all of it run at DRAM speed.  Wide states are a whole
// 32-byte sector per piece, so reading them through a permutation costs nothing extra: here the
// pieces are walked in the order of their START breakpoint.  Lanes are columns; the contributions of
// the pieces that start at one breakpoint (11 on average) are summed in registers and leave as one
// reduction per breakpoint, and the reductions at the END breakpoints land a piece length ahead of
// the walk -- a moving front of a few tens of MB that stays in L2.
constexpr uint32_t BYPOS_CHUNK = 256;  // pieces per warp

__global__ void k_so_keys(uint32_t npp, const uint32_t *__restrict__ q_bp0, const uint32_t *__restrict__ q_bp1,
    uint32_t *key, uint32_t *val, uint32_t pad_key) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npp) return;
    key[j] = q_bp1[j] != NO_PIECE ? q_bp0[j] : pad_key;  // padding sorts last
    val[j] = j;
}

__global__ void k_so_gather(uint32_t n, const uint32_t *__restrict__ slot, const uint32_t *__restrict__ q_bp0,
    const uint32_t *__restrict__ q_bp1, const double *__restrict__ q_bl, uint32_t *bp0, uint32_t *bp1, double *bl) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = slot[i];
    bp0[i] = q_bp0[j]; bp1[i] = q_bp1[j]; bl[i] = q_bl[j];
}


// A warp takes BYPOS_CHUNK consecutive pieces of the summary order, 32 at a time through shared memory
// (state, branch length, breakpoints).  The lanes are split into 32 / CP groups of CP >= ncols lanes
// (CP a power of two): group g walks the g-th part of the 32 pieces, lane c of the group evaluates
// column c.  No shuffles: every lane reads what it needs from the warp's tile.
template <int STAT, class V>
__global__ void __launch_bounds__(TB) k_branch_summary_bypos(uint32_t nsp,
    const uint32_t *__restrict__ so_slot, const uint32_t *__restrict__ so_bp0,
    const uint32_t *__restrict__ so_bp1, const double *__restrict__ so_bl, const V *__restrict__ pval,
    SumP sp, V totals, DeltaOut out, uint32_t m0, uint32_t ncols, uint32_t CP) {
    using T = typename V::scalar;
    using V4 = SVec<T, 4>;
    __shared__ T s_state[TB / 32][32][V::N];
    __shared__ double s_bl[TB / 32][32];
    __shared__ uint32_t s_bp0[TB / 32][32], s_bp1[TB / 32][32];
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t c0 = warp * BYPOS_CHUNK;
    if (c0 >= nsp) return;
    const uint32_t c1 = min(nsp, c0 + BYPOS_CHUNK);
    const uint32_t groups = 32u / CP, per_group = 32u / groups;  // pieces of a tile each group walks
    const uint32_t g = lane / CP, cl = lane % CP;
    const bool mine = cl < ncols;
    const uint32_t m = m0 + (mine ? cl : 0);
    ColP col = out.cols[m];
    V4 tot4 = ivec_zero<V4>();
    int idx4[4] = { col.i, col.j, col.k, col.l };
    if constexpr (stat_reads_tuple_only<STAT>()) {
        // the tuple's sets renumbered 0..3 (one-way statistics: the column's own set)
        if (STAT == STAT_DIVERSITY || STAT == STAT_SEGSITES || STAT == STAT_Y1) idx4[0] = col.i;
#pragma unroll
        for (int a = 0; a < 4; a++) {
            idx4[a] = idx4[a] < 0 || idx4[a] >= V::N ? 0 : idx4[a];
            tot4.v[a] = pick_t<V>(totals, idx4[a]);
        }
        col.i = 0; col.j = 1; col.k = 2; col.l = 3;
    }
    double *Dl = out.D + cl;  // D[bp * ncols + column]
    double acc = 0.0;
    uint32_t cur = NO_PIECE;    // breakpoint the register accumulator belongs to
    for (uint32_t base = c0; base < c1; base += 32) {
        const uint32_t j = base + lane;
        __syncwarp();
        if (j < c1) {
            const V st = pval[so_slot[j]];
#pragma unroll
            for (int k = 0; k < V::N; k++) s_state[wib][lane][k] = st.v[k];
            s_bl[wib][lane] = so_bl[j]; s_bp0[wib][lane] = so_bp0[j]; s_bp1[wib][lane] = so_bp1[j];
        } else {
            s_bp1[wib][lane] = NO_PIECE;
        }
        __syncwarp();
        for (uint32_t q = 0; q < per_group; q++) {
            const uint32_t i = g * per_group + q;
            const uint32_t b1 = s_bp1[wib][i];
            if (b1 == NO_PIECE) break;  // past the end of the chunk (group-uniform)
            const uint32_t b0 = s_bp0[wib][i];
            const double bl_i = s_bl[wib][i];
            double G = 0.0;
            if (!(sp.skip_zero_bl && bl_i == 0.0)) {
                if constexpr (stat_reads_tuple_only<STAT>()) {
                    V4 t4;
#pragma unroll
                    for (int a = 0; a < 4; a++) t4.v[a] = s_state[wib][i][idx4[a]];
                    G = bl_i * F_branch<STAT, V4>(sp, col, (int) m, t4, tot4);
                } else {
                    V s_i;
#pragma unroll
                    for (int k = 0; k < V::N; k++) s_i.v[k] = s_state[wib][i][k];
                    G = bl_i * F_branch<STAT, V>(sp, col, (int) m, s_i, totals);
                }
            }
            if (b0 != cur) {
                if (cur != NO_PIECE && mine && acc != 0.0) atomicAdd(Dl + (size_t) cur * ncols, acc);
                cur = b0;
                acc = 0.0;
            }
            acc += G;
            if (mine && G != 0.0) atomicAdd(Dl + (size_t) b1 * ncols, -G);
        }
    }
    if (cur != NO_PIECE && mine && acc != 0.0) atomicAdd(Dl + (size_t) cur * ncols, acc);
}

// ---- many result columns, window runs (finite summaries, windows x columns small enough; default)
// The same walk by start position, with the window sums kept in registers instead of per-breakpoint
// deltas (cf. k_branch_summary_runs): the pieces arrive sorted by the position they start at, so the
// window of a lane group's register sum changes only when the walk passes a window edge, and a piece
// that ends in a later window (31 % on C2 with 1000 windows) sends its last part there as ONE reduction
// per column -- a third of the reductions of the delta formulation, into bins that stay in L2, and no
// scan or integration over the breakpoints afterwards.  The lane that loads a piece also finds its
// windows (once per piece, not per column).  Concurrent warps take chunks far apart along the genome
// (chunk = warp x chunk_mul mod chunks) and the bins are replicated: few reductions meet at one address.
constexpr uint32_t RUN_DEAD = 0xfffffffeu;  // a piece that adds nothing (no branch above it, or outside the windows)

__device__ __forceinline__ void run_piece_windows(const RunArgs &b, int exact, double x0, double x1, uint32_t &w0,
    uint32_t &w1, double &d0, double &d1) {
    const double a = fmax(x0, b.w0), e = fmin(x1, b.wlast);
    if (!(a < e)) {
        w0 = RUN_DEAD; w1 = RUN_DEAD; d0 = 0.0; d1 = 0.0;
        return;
    }
    double hi0, lo1;
    if (exact) {
        w0 = min((uint32_t) ((a - b.w0) * b.inv_step), b.W - 1);
        w1 = min((uint32_t) ((e - b.w0) * b.inv_step), b.W - 1);
        hi0 = w0 + 1 >= b.W ? b.wlast : b.w0 + (double) (w0 + 1) * b.step;
        lo1 = b.w0 + (double) w1 * b.step;
    } else {
        w0 = run_window_of<false>(b.windows, b, a);
        w1 = run_window_of<true>(b.windows, b, e);
        hi0 = b.windows[w0 + 1];
        lo1 = b.windows[w1];
    }
    d0 = w1 == w0 ? e - a : hi0 - a;
    d1 = w1 == w0 ? 0.0 : e - lo1;
}

template <int STAT, class V>
__global__ void __launch_bounds__(TB) k_branch_summary_bypos_runs(uint32_t nsp,
    const uint32_t *__restrict__ so_slot, const uint32_t *__restrict__ so_bp0,
    const uint32_t *__restrict__ so_bp1, const double *__restrict__ so_bl, const double *__restrict__ bp_pos,
    const V *__restrict__ pval, SumP sp, V totals, const ColP *cols, uint32_t m0, uint32_t ncols, uint32_t CP,
    RunArgs b, int exact, uint32_t nchunks) {
    using T = typename V::scalar;
    using V4 = SVec<T, 4>;
    __shared__ T s_state[TB / 32][32][V::N];
    __shared__ double s_bl[TB / 32][32], s_d0[TB / 32][32], s_d1[TB / 32][32];
    __shared__ uint32_t s_w0[TB / 32][32], s_w1[TB / 32][32];
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= nchunks) return;
    const uint32_t chunk = (uint32_t) (((uint64_t) warp * b.chunk_mul) % nchunks);
    const uint32_t c0 = chunk * BYPOS_CHUNK;
    const uint32_t c1 = min(nsp, c0 + BYPOS_CHUNK);
    const uint32_t groups = 32u / CP, per_group = 32u / groups;  // pieces of a tile each group walks
    const uint32_t g = lane / CP, cl = lane % CP;
    const bool mine = cl < ncols;
    const uint32_t m = m0 + (mine ? cl : 0);
    ColP col = cols[m];
    V4 tot4 = ivec_zero<V4>();
    int idx4[4] = { col.i, col.j, col.k, col.l };
    if constexpr (stat_reads_tuple_only<STAT>()) {
        if (STAT == STAT_DIVERSITY || STAT == STAT_SEGSITES || STAT == STAT_Y1) idx4[0] = col.i;
#pragma unroll
        for (int a = 0; a < 4; a++) {
            idx4[a] = idx4[a] < 0 || idx4[a] >= V::N ? 0 : idx4[a];
            tot4.v[a] = pick_t<V>(totals, idx4[a]);
        }
        col.i = 0; col.j = 1; col.k = 2; col.l = 3;
    }
    const size_t block = (2 * (size_t) b.Wl + 1) * ncols;
    double *gR = b.bins + (size_t) (warp & b.copy_mask) * block + cl - (size_t) b.wlo * ncols;   // R[w * ncols + column]
    double *gC = gR + (size_t) b.Wl * ncols;                                                      // C[w * ncols + column]
    double acc = 0.0;
    uint32_t wc = RUN_DEAD;    // window the register sum belongs to (group-uniform)
    for (uint32_t base = c0; base < c1; base += 32) {
        const uint32_t j = base + lane;
        __syncwarp();
        if (j < c1) {
            const V st = pval[so_slot[j]];
#pragma unroll
            for (int k = 0; k < V::N; k++) s_state[wib][lane][k] = st.v[k];
            const double bl_j = so_bl[j];
            uint32_t w0, w1;
            double d0, d1;
            run_piece_windows(b, exact, bp_pos[so_bp0[j]], bp_pos[so_bp1[j]], w0, w1, d0, d1);
            if (bl_j == 0.0) w0 = RUN_DEAD;  // finite summaries only here: 0 x f adds nothing
            s_bl[wib][lane] = bl_j; s_w0[wib][lane] = w0; s_w1[wib][lane] = w1;
            s_d0[wib][lane] = d0; s_d1[wib][lane] = d1;
        } else {
            s_w0[wib][lane] = NO_PIECE;
        }
        __syncwarp();
        for (uint32_t q = 0; q < per_group; q++) {
            const uint32_t i = g * per_group + q;
            const uint32_t w0 = s_w0[wib][i];
            if (w0 == NO_PIECE) break;  // past the end of the chunk (group-uniform)
            if (w0 == RUN_DEAD) continue;
            double G;
            if constexpr (stat_reads_tuple_only<STAT>()) {
                V4 t4;
#pragma unroll
                for (int a = 0; a < 4; a++) t4.v[a] = s_state[wib][i][idx4[a]];
                G = s_bl[wib][i] * F_branch<STAT, V4>(sp, col, (int) m, t4, tot4);
            } else {
                V s_i;
#pragma unroll
                for (int k = 0; k < V::N; k++) s_i.v[k] = s_state[wib][i][k];
                G = s_bl[wib][i] * F_branch<STAT, V>(sp, col, (int) m, s_i, totals);
            }
            if (w0 != wc) {
                if (wc != RUN_DEAD && mine && acc != 0.0) atomicAdd(gR + (size_t) wc * ncols, acc);
                acc = 0.0;
                wc = w0;
            }
            acc += G * s_d0[wib][i];
            const uint32_t w1 = s_w1[wib][i];
            if (w1 != w0 && mine && G != 0.0) {
                atomicAdd(gR + (size_t) w1 * ncols, G * s_d1[wib][i]);
                if (w1 > w0 + 1) {  // the windows in between are covered entirely
                    atomicAdd(gC + (size_t) (w0 + 1) * ncols, G);
                    atomicAdd(gC + (size_t) w1 * ncols, -G);
                }
            }
        }
    }
    if (wc != RUN_DEAD && mine && acc != 0.0) atomicAdd(gR + (size_t) wc * ncols, acc);
}

// ---- running sums and window integrals of deltas laid out [breakpoint][column], lanes = columns:
// a warp owns COLSCAN_ROWS consecutive breakpoints.  Pass 1: column sums of every chunk; pass 2: their
// exclusive prefix over the chunks (one warp); pass 3: the warp walks its breakpoints again with the
// running sum S and adds S x (overlap of [t_i, t_i+1) with the window) to the windows it meets
// (trees.c:1484-1504) -- S is never written back.
constexpr uint32_t COLSCAN_ROWS = 512;

__global__ void __launch_bounds__(TB) k_colscan_partial(const double *__restrict__ D, uint32_t T, uint32_t ncols,
    double *partial) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t r0 = warp * COLSCAN_ROWS;
    if (r0 >= T || lane >= ncols) return;
    const uint32_t r1 = min(T, r0 + COLSCAN_ROWS);
    const double *p = D + (size_t) r0 * ncols + lane;
    double sum = 0.0;
    uint32_t r = r0;
    for (; r + 8 <= r1; r += 8, p += (size_t) 8 * ncols) {
        double v[8];
#pragma unroll
        for (int q = 0; q < 8; q++) v[q] = p[(size_t) q * ncols];  // eight loads in flight
#pragma unroll
        for (int q = 0; q < 8; q++) sum += v[q];                  // summed in breakpoint order
    }
    for (; r < r1; r++, p += ncols) sum += p[0];
    partial[(size_t) warp * ncols + lane] = sum;
}

// exclusive prefix of the chunk sums over the chunks, per column; one block of 32 warps, lanes = columns
__global__ void __launch_bounds__(1024) k_colscan_offsets(double *partial, uint32_t nchunks, uint32_t ncols) {
    __shared__ double tot[32][33];
    const uint32_t lane = threadIdx.x & 31u, wp = threadIdx.x >> 5;
    const uint32_t per = (nchunks + 31) / 32;
    const uint32_t c0 = min(nchunks, wp * per), c1 = min(nchunks, c0 + per);
    double sum = 0.0;
    if (lane < ncols) {
#pragma unroll 8
        for (uint32_t c = c0; c < c1; c++) sum += partial[(size_t) c * ncols + lane];
    }
    tot[wp][lane] = sum;
    __syncthreads();
    double run = 0.0;
    for (uint32_t q = 0; q < wp; q++) run += tot[q][lane];
    if (lane < ncols) {
#pragma unroll 8
        for (uint32_t c = c0; c < c1; c++) {
            const double v = partial[(size_t) c * ncols + lane];
            partial[(size_t) c * ncols + lane] = run;
            run += v;
        }
    }
}

__global__ void __launch_bounds__(TB) k_colscan_integrate(const double *__restrict__ D, uint32_t T,
    uint32_t ncols, const double *__restrict__ partial, const double *__restrict__ bp_pos,
    const double *__restrict__ windows, uint32_t W, uint32_t m0, uint32_t M, double *result) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t r0 = warp * COLSCAN_ROWS;
    if (r0 >= T) return;
    const uint32_t r1 = min(T, r0 + COLSCAN_ROWS);
    const bool mine = lane < ncols;
    double S = mine ? partial[(size_t) warp * ncols + lane] : 0.0;
    // window holding the first interval's left end (warp-uniform)
    double a = bp_pos[r0];
    uint32_t w = upper_bound_dev(windows, W + 1, a);
    w = w > 0 ? w - 1 : 0;
    if (w >= W) return;  // the chunk starts at or beyond the last window edge
    double wl = windows[w], wr = windows[w + 1];
    double acc = 0.0;
    const double *p = D + (size_t) r0 * ncols + (mine ? lane : 0);
    for (uint32_t r = r0; r < r1; r++, p += ncols) {
        if (mine) S += p[0];
        const double b = bp_pos[r + 1];
        // the interval [a, b) against the windows it meets
        while (true) {
            const double lo = a > wl ? a : wl, hi = b < wr ? b : wr;
            if (hi > lo) acc += S * (hi - lo);
            if (b < wr) break;  // the window goes on past this interval
            if (mine && acc != 0.0) atomicAdd(result + (size_t) w * M + m0 + lane, acc);
            acc = 0.0;
            w++;
            if (w >= W) return;
            wl = wr;
            wr = windows[w + 1];
        }
        a = b;
    }
    if (mine && acc != 0.0) atomicAdd(result + (size_t) w * M + m0 + lane, acc);
}

__global__ void k_span_divide(const double *__restrict__ windows, uint32_t W, uint32_t M, uint32_t m0,
    uint32_t ncols, double *result) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t) W * ncols) return;
    const uint32_t w = (uint32_t) (i / ncols), c = (uint32_t) (i % ncols);
    result[(size_t) w * M + m0 + c] /= windows[w + 1] - windows[w];
}

// ---------------------------------------------------------------- phase 3, branch mode
// S = inclusive prefix sum of D over the breakpoints (the reference's running sum after the diffs
// of breakpoint i, cub::DeviceScan per column); window w gets the integral of S over it:
// sum over the breakpoint intervals [t_i, t_i+1) that meet it of S_i * overlap
// (trees.c:1484-1504), span-normalised (trees.c:1920-1934).  G threads per (window, column):
// a warp, or a whole block when there are few windows.
// first index in [0, n) with a[idx] > x (UPPER) or a[idx] >= x (!UPPER), searched by the whole block:
// every round narrows the range by a factor of blockDim.x with one load per thread (3 dependent
// loads for 4.4 M breakpoints instead of 22).  Must be called by all threads of the block.
template <bool UPPER>
__device__ __forceinline__ uint32_t block_bound(const double *__restrict__ a, uint32_t n, double x) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > blockDim.x) {
        const uint32_t step = (hi - lo + blockDim.x - 1) / blockDim.x;
        const uint32_t idx = lo + threadIdx.x * step;
        bool below = false;  // a[idx] is still left of the bound
        if (idx < hi) below = UPPER ? !(a[idx] > x) : !(a[idx] >= x);
        const uint32_t f = (uint32_t) __syncthreads_count(below);  // samples are monotone: a prefix is below
        if (f == 0) return lo;
        const uint32_t nlo = lo + (f - 1) * step + 1;
        hi = min(hi, lo + f * step);
        lo = nlo;
    }
    bool below = false;
    if (lo + threadIdx.x < hi) below = UPPER ? !(a[lo + threadIdx.x] > x) : !(a[lo + threadIdx.x] >= x);
    return lo + (uint32_t) __syncthreads_count(below);
}

template <int G>
__global__ void __launch_bounds__(TB) k_window_integrate(const double *S, uint32_t Tp1,
    const double *__restrict__ bp_pos, const double *__restrict__ windows, uint32_t W, uint32_t mcols,
    uint32_t m0, uint32_t M, int span_normalise, double *result) {
    __shared__ double s_part[TB / 32];
    const uint32_t T = Tp1 - 1;
    const uint32_t lane = threadIdx.x & 31u;
    const size_t group = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const uint32_t gt = threadIdx.x % G;
    const bool active = group < (size_t) W * mcols;  // uniform over the group
    double acc = 0.0, wl = 0.0, wr = 1.0;
    uint32_t w = 0, mloc = 0;
    if (active) {
        w = (uint32_t) (group / mcols);
        mloc = (uint32_t) (group % mcols);
        wl = windows[w];
        wr = windows[w + 1];
        const double *Sm = S + (size_t) mloc * Tp1;
        // intervals meeting [wl, wr): from the one holding wl (or the first) to the last starting < wr
        uint32_t lo, hi;
        if (G == TB) {  // one window per block: the block searches together (active is block-uniform)
            lo = block_bound<true>(bp_pos, T, wl);
            hi = block_bound<false>(bp_pos, T, wr);
        } else {
            lo = upper_bound_dev(bp_pos, T, wl);
            hi = lower_bound_dev(bp_pos, T, wr);
        }
        lo = lo > 0 ? lo - 1 : 0;
        for (uint32_t i = lo + gt; i < hi; i += G) {
            double a = bp_pos[i], b = bp_pos[i + 1];
            a = a > wl ? a : wl;
            b = b < wr ? b : wr;
            if (b > a) acc += Sm[i] * (b - a);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
    if (G > 32) {
        if (lane == 0) s_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            acc = 0.0;
            for (int i = 0; i < TB / 32; i++) acc += s_part[i];
        }
    }
    if (active && gt == 0) {
        if (span_normalise) acc /= wr - wl;
        result[(size_t) w * M + m0 + mloc] = acc;
    }
}

// ---------------------------------------------------------------- phase 2, site mode

template <int STAT, class V>
__global__ void k_site_summary(uint32_t site_lo, uint32_t nsites, const uint32_t *site_moff,
    const uint32_t *site_aoff, const int32_t *mut_src, const uint16_t *mut_allele,
    const uint16_t *mut_alt, const V *pval, V totals, V *scratch, SumP P,
    double *R) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nsites) return;
    uint32_t site = site_lo + t;
    uint32_t a0 = site_aoff[site], na = site_aoff[site + 1] - a0;
    const uint32_t mb = site_moff[site], me = site_moff[site + 1];
    if (na == 2 && me - mb == 1) {
        // one mutation, two alleles: no scratch traffic
        V x = pval[mut_src[mb]];
        V anc = totals - x;
        for (int m = 0; m < P.M; m++) {
            const ColP col = P.cols[m];
            double acc = 0.0;
            if (!P.polarised) acc += f_eval<STAT, V>(P, col, m, anc);
            acc += f_eval<STAT, V>(P, col, m, x);
            R[(size_t) m * nsites + t] = acc;
        }
        return;
    }
    scratch[a0] = totals;  // allele 0 starts at total_weight (trees.c:1548)
    for (uint32_t al = 1; al < na; al++) scratch[a0 + al] = ivec_zero<V>();
    for (uint32_t m = mb; m < me; m++) {
        V x = pval[mut_src[m]];
        scratch[a0 + mut_allele[m]] = scratch[a0 + mut_allele[m]] + x;
        scratch[a0 + mut_alt[m]] = scratch[a0 + mut_alt[m]] - x;
    }
    for (int m = 0; m < P.M; m++) {
        const ColP col = P.cols[m];
        double acc = 0.0;
        for (uint32_t al = P.polarised ? 1 : 0; al < na; al++) {
            acc += f_eval<STAT, V>(P, col, m, scratch[a0 + al]);
        }
        R[(size_t) m * nsites + t] = acc;
    }
}

// site: result[w] = sum of site results with windows[w] <= position < windows[w+1]  (trees.c:1753-1762)
__global__ void k_window_site(const double *windows, uint32_t nsplit, const double *site_pos,
    uint32_t site_lo, uint32_t nsites, const double *R, uint32_t M, double *partial) {
    typedef cub::BlockReduce<double, TB> BR;
    __shared__ typename BR::TempStorage tmp;
    uint32_t w = blockIdx.x, c = blockIdx.y;
    double wl = windows[w], wr = windows[w + 1];
    const double *pos = site_pos + site_lo;
    uint32_t lo = lower_bound_dev(pos, nsites, wl);
    uint32_t hi = lower_bound_dev(pos, nsites, wr);
    uint32_t len = hi - lo, per = (len + nsplit - 1) / nsplit;
    uint32_t s0 = lo + c * per, s1 = s0 + per;
    if (s0 > hi) s0 = hi;
    if (s1 > hi) s1 = hi;
    for (uint32_t m = 0; m < M; m++) {
        const double *row = R + (size_t) m * nsites;
        double sum = 0.0;
        for (uint32_t i = s0 + threadIdx.x; i < s1; i += TB) sum += row[i];
        double tot = BR(tmp).Sum(sum);
        if (threadIdx.x == 0) partial[((size_t) w * nsplit + c) * M + m] = tot;
        __syncthreads();
    }
}

__global__ void k_window_final(const double *partial, const double *windows, uint32_t W,
    uint32_t nsplit, uint32_t M, int span_normalise, double *result) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= W * M) return;
    uint32_t w = t / M, m = t % M;
    double sum = 0.0;
    for (uint32_t c = 0; c < nsplit; c++) sum += partial[((size_t) w * nsplit + c) * M + m];
    if (span_normalise) sum /= windows[w + 1] - windows[w];  // trees.c:1920-1934
    result[t] = sum;
}

// ---------------------------------------------------------------- allele frequency spectrum, site mode
// tsk_treeseq_update_site_afs (trees.c:3497-3540): for every allele of a site carried by some but
// not all samples, +1 (polarised; the ancestral allele is skipped) or +1/2 (folded) at the vector
// of per-set allele counts in the window holding the site.  State columns: the K sample sets and,
// last, all samples.

// fold (trees.c:3469-3495): the lexicographically smaller of a coordinate and its mirror image
__device__ __forceinline__ void afs_fold(uint32_t *coord, const uint32_t *dims, int K) {
    double n = 0;
    int s = 0;
    for (int k = 0; k < K; k++) {
        n += (double) dims[k] - 1;
        s += (int) coord[k];
    }
    n /= 2;
    int k = K;
    while (s == n && k > 0) {
        k--;
        n -= ((double) (dims[k] - 1)) / 2;
        s -= (int) coord[k];
    }
    if (s > n) {
        for (k = 0; k < K; k++) coord[k] = dims[k] - 1 - coord[k];
    }
}

// Up to 7 sample sets: one int32 state column per set, then the all-samples column.  More sets
// (packed): the spectrum's row-major coordinate is itself a linear function of the per-set counts, so
// it is carried as ONE state column (fp64, exact) next to the all-samples count.
constexpr int AFS_MAX_PACKED_SETS = 64;
struct AfsShape {
    int packed;
    uint32_t nsets;
    const uint32_t *dims;  // device [nsets], packed only
};

// row-major index (increment_nd_array_value) of the state's coordinate, folded unless polarised; false
// when the allele / branch is carried by no sample or by all (trees.c:3519, 3622)
template <class V>
__device__ __forceinline__ bool afs_index(const V &cnt, const SumP &P, const AfsShape &sh, uint32_t num_samples,
    size_t &index) {
    if (!sh.packed) {
        const int K = P.K - 1;
        uint32_t coord[8], dims[8], total = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            coord[k] = 0;
            dims[k] = 1;
        }
#pragma unroll
        for (int k = 0; k < V::N; k++) {
            if (k < K) {
                coord[k] = (uint32_t) cnt.v[k];
                dims[k] = (uint32_t) P.n[k] + 1;
            }
            if (k == K) total = (uint32_t) cnt.v[k];
        }
        if (!(total > 0 && total < num_samples)) return false;
        if (!P.polarised) afs_fold(coord, dims, K);
        index = 0;
        for (int k = 0; k < K; k++) index = index * dims[k] + coord[k];
        return true;
    }
    unsigned long long lin = (unsigned long long) cnt.v[0];
    const uint32_t total = (uint32_t) cnt.v[V::N > 1 ? 1 : 0];
    if (!(total > 0 && total < num_samples)) return false;
    if (!P.polarised) {
        uint32_t coord[AFS_MAX_PACKED_SETS];
        const int K = (int) sh.nsets;
        for (int k = K - 1; k >= 0; k--) {
            coord[k] = (uint32_t) (lin % sh.dims[k]);
            lin /= sh.dims[k];
        }
        afs_fold(coord, sh.dims, K);
        lin = 0;
        for (int k = 0; k < K; k++) lin = lin * sh.dims[k] + coord[k];
    }
    index = (size_t) lin;
    return true;
}

template <class V>
__global__ void k_site_afs(uint32_t site_lo, uint32_t nsites, const uint32_t *site_moff,
    const uint32_t *site_aoff, const int32_t *mut_src, const uint16_t *mut_allele, const uint16_t *mut_alt,
    const V *pval, V totals, V *scratch, SumP P, AfsShape sh, uint32_t num_samples, const double *site_pos,
    const double *windows, uint32_t W, size_t afs_size, double *result) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nsites) return;
    const uint32_t site = site_lo + t;
    uint32_t w = upper_bound_dev(windows, W + 1, site_pos[site]);
    w = w > 0 ? w - 1 : 0;
    if (w >= W) w = W - 1;
    double *afs = result + (size_t) w * afs_size;
    const double inc = P.polarised ? 1.0 : 0.5;
    const uint32_t a0 = site_aoff[site], na = site_aoff[site + 1] - a0;
    const uint32_t mb = site_moff[site], me = site_moff[site + 1];
    scratch[a0] = totals;  // allele 0 starts at the totals (trees.c:1548)
    for (uint32_t al = 1; al < na; al++) scratch[a0 + al] = ivec_zero<V>();
    for (uint32_t m = mb; m < me; m++) {
        V x = pval[mut_src[m]];
        scratch[a0 + mut_allele[m]] = scratch[a0 + mut_allele[m]] + x;
        scratch[a0 + mut_alt[m]] = scratch[a0 + mut_alt[m]] - x;
    }
    for (uint32_t al = P.polarised ? 1 : 0; al < na; al++) {
        size_t index;
        if (afs_index<V>(scratch[a0 + al], P, sh, num_samples, index)) atomicAdd(afs + index, inc);
    }
}

// Branch mode (tsk_treeseq_branch_allele_frequency_spectrum, trees.c:3699-3812): every node with a
// parent adds (span) x (branch length inside each time window) at the vector of per-set sample
// counts below it.  The span of a piece runs from the node's last update -- q_eff, never before
// the left edge of the window holding the piece's start, because every window end flushes all
// nodes -- to the piece's end, split over the windows it crosses.  Time windows (trees.c:3663-3680):
// window k takes max(0, min(tw[k+1], t_v) - max(tw[k], t_u)) of the branch from t_u up to t_v, for
// every k with tw[k] < t_v; q_node == nullptr: the default window [0, inf) with node times >= 0, where
// that is the whole branch.
template <class V>
__global__ void k_branch_afs(uint32_t npp, const uint32_t *__restrict__ q_bp0,
    const uint32_t *__restrict__ q_bp1, const uint32_t *__restrict__ q_eff,
    const double *__restrict__ q_bl, const double *__restrict__ bp_pos, const V *__restrict__ pval, SumP P,
    AfsShape sh, uint32_t num_samples, double range_left, const double *__restrict__ windows, uint32_t W,
    size_t afs_size, const int32_t *__restrict__ q_node, const double *__restrict__ time,
    const double *__restrict__ tw, uint32_t NTW, double *result) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npp) return;
    const uint32_t b1 = q_bp1[j];
    const double bl = q_bl[j];
    if (b1 == NO_PIECE || bl == 0.0) return;
    size_t index;
    if (!afs_index<V>(pval[j], P, sh, num_samples, index)) return;
    const double x = bp_pos[q_bp0[j]], xe = bp_pos[b1];
    const uint32_t e = q_eff[j];
    double start = e == NO_PIECE ? range_left : bp_pos[e];
    uint32_t w = upper_bound_dev(windows, W + 1, x);
    w = w > 0 ? w - 1 : 0;
    if (start < windows[w]) start = windows[w];
    double t_u = 0.0, t_v = 0.0;
    if (q_node != nullptr) {
        t_u = time[q_node[j]];
        t_v = t_u + bl;  // bl = time[parent] - time[node]
    }
    for (; w < W && windows[w] < xe; w++) {
        const double wl = windows[w], wr = windows[w + 1];
        const double len = (xe < wr ? xe : wr) - (start > wl ? start : wl);
        if (!(len > 0.0)) continue;
        if (q_node == nullptr) {
            atomicAdd(result + (size_t) w * afs_size + index, len * bl);
            continue;
        }
        for (uint32_t k = 0; k < NTW && tw[k] < t_v; k++) {
            const double hi = tw[k + 1] < t_v ? tw[k + 1] : t_v, lo = tw[k] > t_u ? tw[k] : t_u;
            const double tbl = hi - lo > 0.0 ? hi - lo : 0.0;
            if (tbl > 0.0) atomicAdd(result + ((size_t) w * NTW + k) * afs_size + index, len * tbl);
        }
    }
}

__global__ void k_afs_span_normalise(const double *windows, uint32_t W, size_t afs_size, double *result) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t) W * afs_size) return;
    const uint32_t w = (uint32_t) (i / afs_size);
    result[i] /= windows[w + 1] - windows[w];
}

// ---------------------------------------------------------------- relatedness vector
// tsk_treeseq_genetic_relatedness_vector (trees.c:10445-10816), branch mode: out[focal][k] is the
// integral over the window of  sum over the ancestors-or-self u of the focal node of
// branch_length[u] * w_u[k],  w_u = the summed weights of the samples below u (the sweep's state).
// The reference carries this as lazily updated per-node v/x vectors along its sequential sweep; here
// it is the transpose of the sweep: G[piece] = branch_length * |piece ^ window| * state[piece] + the
// G of every piece that references it (a parent piece lies inside the span of each piece it
// references, see plan.cuh), pushed down one height per launch, tallest first.  A sample's INIT slot
// collects the G of all its pieces = its output row.  Focal nodes that are not samples need the
// node of every piece (node-mode plan): their pieces are summed per node.
template <class V>
__global__ void k_relvec_push(uint32_t lo, uint32_t count, const uint32_t *__restrict__ q_bp0,
    const uint32_t *__restrict__ q_bp1, const double *__restrict__ q_bl, const double *__restrict__ bp_pos,
    const uint32_t *__restrict__ q_off, const uint32_t *__restrict__ refs, const V *__restrict__ pval,
    V *G, size_t slots, const double *__restrict__ windows, uint32_t W, uint32_t w0, uint32_t wc,
    const int32_t *__restrict__ q_node, double *Gn, size_t gn_stride, uint32_t K) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const uint32_t j = lo + t;
    const uint32_t b1 = q_bp1[j];
    if (b1 == NO_PIECE) return;
    // G of window w is zero unless the piece meets w (the pieces that reference it lie inside its
    // span): only those windows of this launch's chunk [w0, w0 + wc) are touched
    const double x0 = bp_pos[q_bp0[j]], x1 = bp_pos[b1];
    uint32_t w = upper_bound_dev(windows, W + 1, x0);
    w = w > 0 ? w - 1 : 0;
    if (w < w0) w = w0;
    const uint32_t wend = w0 + wc;
    if (w >= wend || !(windows[w] < x1)) return;
    const double bl = q_bl[j];
    const V st = pval[j];
    const uint32_t r0 = q_off[j], r1 = q_off[j + 1];
    for (; w < wend && windows[w] < x1; w++) {
        V *Gw = G + (size_t) (w - w0) * slots;
        V g = Gw[j];  // complete: every piece referencing j is taller and was pushed by an earlier launch
        const double wl = windows[w], wr = windows[w + 1];
        const double len = (x1 < wr ? x1 : wr) - (x0 > wl ? x0 : wl);
        if (bl != 0.0 && len > 0.0) {
            const double area = bl * len;
#pragma unroll
            for (int k = 0; k < V::N; k++) g.v[k] += area * st.v[k];
        }
        bool any = false;
#pragma unroll
        for (int k = 0; k < V::N; k++) any |= g.v[k] != 0.0;
        if (!any) continue;
        if (Gn != nullptr) {
            double *dst = Gn + (size_t) (w - w0) * gn_stride + (size_t) q_node[j] * K;
#pragma unroll
            for (int k = 0; k < V::N; k++) {
                if ((uint32_t) k < K) atomicAdd(dst + k, g.v[k]);
            }
        }
        for (uint32_t r = r0; r < r1; r++) {
            double *dst = reinterpret_cast<double *>(Gw + refs[r]);
#pragma unroll
            for (int k = 0; k < V::N; k++) {
                if ((uint32_t) k < K) atomicAdd(dst + k, g.v[k]);
            }
        }
    }
}

// rows of the focal nodes, for the wc windows of one chunk: out[c][j][k]
template <class V>
__global__ void k_relvec_out(const int32_t *__restrict__ focal, uint32_t nf, uint32_t K, uint32_t wc,
    const int32_t *__restrict__ sample_index, const V *__restrict__ G, size_t slots, uint32_t npp,
    const double *__restrict__ Gn, size_t gn_stride, double *out) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t row = (size_t) nf * K;
    if (i >= row * wc) return;
    const uint32_t c = (uint32_t) (i / row), j = (uint32_t) ((i % row) / K), k = (uint32_t) (i % K);
    const int32_t u = focal[j];
    const int32_t si = sample_index[u];
    double r = 0.0;
    if (si >= 0) {
        r = reinterpret_cast<const double *>(G + (size_t) c * slots + npp + si)[k];
    } else if (Gn != nullptr) {
        r = Gn[(size_t) c * gn_stride + (size_t) u * K + k];
    }
    out[i] = r;
}

// ---------------------------------------------------------------- node mode
// tsk_treeseq_node_general_stat (trees.c:1788-1918): result[w][u] is the integral over window w of
// the summary of node u's state -- no branch lengths, and every node counts, in a tree or not.
// A node's state is constant over each of its pieces and equal to its own weight before its first
// piece; the plan keeps every piece for this (TSKB_INIT_NODE_MODE).

template <int STAT, class V>
__device__ __forceinline__ void node_add(const SumP &sp, const V &totals, const V &st, int32_t node,
    double x, double xe, const double *__restrict__ windows, uint32_t W, uint32_t N, int span_normalise,
    double *result) {
    if (!(xe > x)) return;
    uint32_t w = upper_bound_dev(windows, W + 1, x);
    w = w > 0 ? w - 1 : 0;
    for (; w < W && windows[w] < xe; w++) {
        const double wl = windows[w], wr = windows[w + 1];
        double len = (xe < wr ? xe : wr) - (x > wl ? x : wl);
        if (!(len > 0.0)) continue;
        if (span_normalise) len /= wr - wl;
        double *row = result + ((size_t) w * N + (size_t) node) * sp.M;
        for (int m = 0; m < sp.M; m++) {
            const double v = F_branch<STAT, V>(sp, sp.cols[m], m, st, totals);
            if (v != 0.0) atomicAdd(row + m, v * len);
        }
    }
}

template <int STAT, class V>
__global__ void k_node_pieces(uint32_t npp, const int32_t *__restrict__ q_node,
    const uint32_t *__restrict__ q_bp0, const uint32_t *__restrict__ q_bp1,
    const double *__restrict__ bp_pos, const V *__restrict__ pval, SumP sp, V totals,
    const double *__restrict__ windows, uint32_t W, uint32_t N, int span_normalise, double *result) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npp) return;
    const int32_t node = q_node[j];
    if (node < 0) return;
    node_add<STAT, V>(sp, totals, pval[j], node, bp_pos[q_bp0[j]], bp_pos[q_bp1[j]], windows, W, N,
        span_normalise, result);
}

template <int STAT, class V>
__global__ void k_node_initial(uint32_t N, const uint32_t *__restrict__ node_first_bp,
    const int32_t *__restrict__ sample_index, const V *__restrict__ init, uint32_t zero_slot,
    const double *__restrict__ bp_pos, double range_left, SumP sp, V totals,
    const double *__restrict__ windows, uint32_t W, int span_normalise, double *result) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= N) return;
    const int32_t si = sample_index[u];
    node_add<STAT, V>(sp, totals, init[si >= 0 ? (uint32_t) si : zero_slot], (int32_t) u, range_left,
        bp_pos[node_first_bp[u]], windows, W, N, span_normalise, result);
}

// ---------------------------------------------------------------- trees_at

__global__ void k_parent_at(const double *positions, uint32_t nq, uint32_t N,
    const uint32_t *coff, const double *csr_left, const double *csr_right,
    const int32_t *csr_parent, int32_t *out_parent) {
    size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t) nq * N) return;
    uint32_t q = (uint32_t) (t / N), u = (uint32_t) (t % N);
    double x = positions[q];
    uint32_t lo = coff[u], hi = coff[u + 1];
    uint32_t k = upper_bound_dev(csr_left + lo, hi - lo, x);  // edges with left <= x
    int32_t p = -1;
    if (k > 0 && csr_right[lo + k - 1] > x) p = csr_parent[lo + k - 1];
    out_parent[t] = p;
}

__global__ void k_count_at(const int32_t *tracked, uint32_t nt, uint32_t nq, uint32_t N,
    const int32_t *parent, int32_t *count) {
    size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t) nq * nt) return;
    uint32_t q = (uint32_t) (t / nt);
    int32_t u = tracked[t % nt];
    const int32_t *par = parent + (size_t) q * N;
    int32_t *cnt = count + (size_t) q * N;
    uint32_t guard = 0;
    while (u != -1 && guard++ < (1u << 24)) {
        atomicAdd(&cnt[u], 1);
        u = par[u];
    }
}

// ---------------------------------------------------------------- driver

constexpr size_t DELTA_BUDGET = size_t(1) << 30;   // bytes of per-breakpoint deltas held at once
constexpr uint32_t COLS_KERNEL_MIN = 6;            // result columns from which lanes-are-columns pays

struct CallCtx {
    const Plan *P;
    const StatSpec *sp;
    cudaStream_t s;
    SumP sumP;
    double *d_windows;
    double *d_result;
    int *d_err;
    uint32_t *d_counters;
    uint64_t launches;
    bool skip_sweep;  // the "states" handed to the summary are already final (run_general_stat)
};

template <class V>
void launch_sweep(CallCtx &c, V *pval) {
    const Plan &P = *c.P;
    Arena &A = P.arena;
    if (P.ntiles == 0 || c.skip_sweep) return;
    uint32_t *counters = c.d_counters;  // zeroed with the per-call staging copy
    unsigned long long *trace = nullptr;
    if (getenv("TSKB_TRACE") != nullptr) {
        trace = A.get<unsigned long long>((size_t) P.ntiles * 4);
        TSKB_CK(cudaMemsetAsync(trace, 0, (size_t) P.ntiles * 4 * sizeof(unsigned long long), c.s));
        P.stats_trace = trace;
    }
    auto kern = k_sweep<V>;
    constexpr int PROP_TB = prop_tb<V>();
    int per_sm = 1, sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, P.device);
    TSKB_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PROP_TB, 0));
    // experiments only (tools/concurrency_probe.py): leave room for another stream's kernels
    if (const char *cap = getenv("TSKB_SWEEP_MAX_PER_SM")) per_sm = std::min(per_sm, std::max(1, atoi(cap)));
    const uint32_t grid = std::min<uint32_t>(P.ntiles, (uint32_t) (sms * std::max(per_sm, 1)));
    SweepArgs a = {};
    a.ntiles = P.ntiles;
    a.tile_dep = P.tile_dep.p; a.q_off = P.q_off.p; a.refs = P.refs.p;
    a.counters = counters; a.error_flag = c.d_err; a.trace = trace;
    void *args[] = { &a, &pval };
    TSKB_CK(cudaLaunchCooperativeKernel((const void *) kern, dim3(grid), dim3(PROP_TB), args, 0, c.s));
    c.launches++;
}

// D (deltas of mcols columns) -> running sums -> window integrals of columns [m0, m0 + mcols)
inline void finish_columns(CallCtx &c, double *D, uint32_t Tp1, uint32_t m0, uint32_t mcols,
    void *scan_tmp, size_t scan_bytes) {
    const Plan &P = *c.P;
    const uint32_t W = c.sp->W, M = c.sp->M;
    for (uint32_t m = 0; m < mcols && Tp1 > 1; m++) {
        double *Dm = D + (size_t) m * Tp1;
        TSKB_CK(cub::DeviceScan::InclusiveSum(scan_tmp, scan_bytes, Dm, Dm, (int) (Tp1 - 1), c.s));
        c.launches++;
    }
    const size_t groups = (size_t) W * mcols;
    const int span = (c.sp->options & TSKB_STAT_SPAN_NORMALISE) ? 1 : 0;
    if (groups * 32 < (size_t) 148 * 2048) {  // few windows: a block per (window, column)
        k_window_integrate<TB><<<(unsigned) groups, TB, 0, c.s>>>(D, Tp1, P.bp_pos.p, c.d_windows, W,
            mcols, m0, M, span, c.d_result);
    } else {
        k_window_integrate<32><<<grid_for(groups * 32, TB), TB, 0, c.s>>>(D, Tp1, P.bp_pos.p,
            c.d_windows, W, mcols, m0, M, span, c.d_result);
    }
    TSKB_CK_LAUNCH();
    c.launches++;
}

// the summary order of the plan (pieces by start breakpoint), built once on first use
__global__ void k_so_count(uint32_t npp, const uint32_t *__restrict__ key_sorted, uint32_t pad_key, uint32_t *nsp) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npp) return;
    const bool real = key_sorted[i] != pad_key;
    if (i == 0 && !real) *nsp = 0;
    if (real && (i + 1 == npp || key_sorted[i + 1] == pad_key)) *nsp = i + 1;
}

void ensure_summary_order(const Plan &P, cudaStream_t s) {
    if (P.so_built) return;
    const uint32_t npp = P.npp;
    P.nsp = 0;
    if (npp > 0) {
        DevArray<uint32_t> key, key_out, val, d_nsp;
        key.alloc(npp); key_out.alloc(npp); val.alloc(npp); d_nsp.alloc(1);
        P.so_slot.alloc(npp);
        const uint32_t pad_key = P.T + 1;
        k_so_keys<<<grid_for(npp, TB), TB, 0, s>>>(npp, P.q_bp0.p, P.q_bp1.p, key.p, val.p, pad_key);
        TSKB_CK_LAUNCH();
        size_t bytes = 0;
        const int bits = (int) std::max(1u, ceil_log2((uint64_t) pad_key + 1));
        TSKB_CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, key.p, key_out.p, val.p, P.so_slot.p, npp, 0, bits, s));
        DevArray<char> tmp;
        tmp.alloc(bytes);
        TSKB_CK(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, key.p, key_out.p, val.p, P.so_slot.p, npp, 0, bits, s));
        TSKB_CK(cudaMemsetAsync(d_nsp.p, 0, sizeof(uint32_t), s));
        k_so_count<<<grid_for(npp, TB), TB, 0, s>>>(npp, key_out.p, pad_key, d_nsp.p);
        TSKB_CK_LAUNCH();
        uint32_t h = 0;
        TSKB_CK(cudaMemcpyAsync(&h, d_nsp.p, sizeof(h), cudaMemcpyDeviceToHost, s));
        TSKB_CK(cudaStreamSynchronize(s));
        P.nsp = h;
        P.so_bp0.alloc(h); P.so_bp1.alloc(h); P.so_bl.alloc(h);
        if (h) {
            k_so_gather<<<grid_for(h, TB), TB, 0, s>>>(h, P.so_slot.p, P.q_bp0.p, P.q_bp1.p, P.q_bl.p,
                P.so_bp0.p, P.so_bp1.p, P.so_bl.p);
            TSKB_CK_LAUNCH();
        }
        TSKB_CK(cudaStreamSynchronize(s));
    }
    P.so_built = true;
    P.stats.device_bytes = P.device_bytes();
}

void ensure_piece_positions(const Plan &P, cudaStream_t s) {
    if (P.qx_built) return;
    P.q_x0.alloc(P.npp);
    P.q_x1.alloc(P.npp);
    if (P.npp) {
        k_piece_positions<<<grid_for(P.npp, TB), TB, 0, s>>>(P.npp, P.q_bp0.p, P.q_bp1.p, P.bp_pos.p, P.q_x0.p,
            P.q_x1.p);
        TSKB_CK_LAUNCH();
        TSKB_CK(cudaStreamSynchronize(s));
    }
    P.qx_built = true;
    P.stats.device_bytes = P.device_bytes();
}

constexpr size_t BINS_SMEM_MAX = 160 * 1024;  // leaves room for more than one CTA of small-W calls per SM

// Window bins in shared memory: finite summaries, and windows x columns small enough.  Returns false
// when the call does not qualify (the delta formulation runs instead).
template <int STAT, class V>
bool run_branch_bins(CallCtx &c, V *pval, V totals) {
    const Plan &P = *c.P;
    const uint32_t M = c.sp->M, W = c.sp->W;
    // Measured on C2 (profiles/r2b_ab_summary.txt): 1.52 ms per step against 0.98 ms for the delta
    // formulation -- the compare-and-swap additions in shared memory (ATOMS.CAST.SPIN) serialise the
    // lanes of a warp that hit one window, which is most of them.  Kept selectable: TSKB_SUM_VARIANT=bins.
    const char *variant = getenv("TSKB_SUM_VARIANT");
    if (variant == nullptr || variant[0] != 'b') return false;
    if (!c.sumP.skip_zero_bl || P.npp == 0 || P.T == 0) return false;
    if ((size_t) (3 * (size_t) W + 1) * sizeof(double) > BINS_SMEM_MAX) return false;
    const uint32_t mc = (uint32_t) std::min<size_t>(M, (BINS_SMEM_MAX / sizeof(double) - W - 1) / (2 * (size_t) W));
    // many columns that do not fit one pass: the walk by start breakpoint (lanes = columns) reads the
    // states once for up to 32 columns
    if (M >= COLS_KERNEL_MIN && mc < M) return false;
    Arena &A = P.arena;
    ensure_piece_positions(P, c.s);
    double *gP = A.get<double>((size_t) 2 * mc * W);
    double *gC = gP + (size_t) mc * W;
    // uniform windows (np.linspace): the window of a position is a multiplication away
    const double *w = c.sp->windows;
    BinArgs b = {};
    b.windows = c.d_windows; b.W = W; b.w0 = w[0]; b.inv_width = (double) W / (w[W] - w[0]); b.uniform = 1;
    for (uint32_t i = 0; i <= W && b.uniform; i++) {
        const double ideal = w[0] + (w[W] - w[0]) * ((double) i / (double) W);
        if (fabs(w[i] - ideal) * b.inv_width > 0.25) b.uniform = 0;
    }
    b.gP = gP; b.gC = gC;
    launch_sweep<V>(c, pval);
    TSKB_CK(cudaEventRecord(P.ev[2], c.s));
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, P.device);
    auto kern = k_branch_summary_bins<STAT, V>;
    TSKB_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) BINS_SMEM_MAX));
    for (uint32_t m0 = 0; m0 < M; m0 += mc) {
        const uint32_t nc = std::min(M, m0 + mc) - m0;
        const size_t smem = ((size_t) W + 1 + 2 * (size_t) nc * W) * sizeof(double);
        TSKB_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SUM_TB, smem));
        int mult = std::max(per_sm, 1);
        if (const char *e = getenv("TSKB_BINS_GRID_MULT")) mult = std::max(1, atoi(e));  // experiments
        const uint32_t ntiles = P.npp / SUM_TB;
        TSKB_CK(cudaMemsetAsync(gP, 0, (size_t) 2 * mc * W * sizeof(double), c.s));
        // gC of this column group must follow its gP: both live in one block of 2 * mc * W doubles
        b.gP = gP; b.gC = gP + (size_t) nc * W;
        kern<<<std::min<uint32_t>(ntiles, (uint32_t) (sms * mult)), SUM_TB, smem, c.s>>>(P.npp, P.q_x0.p, P.q_x1.p,
            P.q_bl.p, pval, c.sumP, totals, c.sumP.cols, m0, nc, b);
        TSKB_CK_LAUNCH();
        c.launches++;
        if (m0 == 0) TSKB_CK(cudaEventRecord(P.ev[3], c.s));
        k_bins_finalize<<<grid_for((size_t) nc * 32, TB), TB, 0, c.s>>>(b.gP, b.gC, c.d_windows, W, nc, m0, M,
            (c.sp->options & TSKB_STAT_SPAN_NORMALISE) ? 1 : 0, c.d_result);
        TSKB_CK_LAUNCH();
        c.launches++;
    }
    TSKB_CK(cudaEventRecord(P.ev[4], c.s));
    return true;
}

// window description for the run kernels; exact = the windows are np.linspace-like: every edge is
// w0 + i * step bit for bit (two roundings, as on the device), the last one the stop value
RunArgs make_run_args(const double *w, uint32_t W, const double *d_windows, bool &exact, double range_left,
    double range_right) {
    RunArgs b = {};
    {
        // windows meeting [range_left, range_right], with one window of margin on either side (the
        // arithmetic window lookup may name a neighbour within a few ulp of an edge)
        const uint32_t i0 = (uint32_t) (std::upper_bound(w, w + W + 1, range_left) - w);   // first edge > left
        const uint32_t i1 = (uint32_t) (std::lower_bound(w, w + W + 1, range_right) - w);  // first edge >= right
        const uint32_t lo = i0 >= 2 ? i0 - 2 : 0;
        const uint32_t hi = std::min<uint32_t>(W, i1 + 1);
        b.wlo = std::min(lo, W - 1);
        b.Wl = std::max<uint32_t>(hi, b.wlo + 1) - b.wlo;
    }
    b.windows = d_windows; b.W = W; b.w0 = w[0]; b.inv_width = (double) W / (w[W] - w[0]); b.uniform = 1;
    for (uint32_t i = 0; i <= W && b.uniform; i++) {
        const double ideal = w[0] + (w[W] - w[0]) * ((double) i / (double) W);
        if (fabs(w[i] - ideal) * b.inv_width > 0.25) b.uniform = 0;
    }
    b.step = W > 1 ? w[1] - w[0] : w[W] - w[0];
    b.inv_step = 1.0 / b.step;
    b.wlast = w[W];
    exact = b.step > 0 && getenv("TSKB_RUN_GENERAL") == nullptr;
    for (uint32_t i = 0; i < W && exact; i++) {
        volatile double prod = (double) i * b.step;
        volatile double e = w[0] + prod;
        if (e != w[i]) exact = false;
    }
    if (exact && !(w[W] > w[0] + (double) (W - 1) * b.step)) exact = false;
    return b;
}

constexpr uint32_t RUNS_MAX_WL = 32768;  // windows met by the engine's range: at least 32 copies of the bins in 16 MB

// Window runs in registers: finite summaries, few columns, windows few enough.  Returns false when the
// call does not qualify (the delta formulation runs instead).
template <int STAT, class V>
bool run_branch_runs(CallCtx &c, V *pval, V totals) {
    const Plan &P = *c.P;
    const uint32_t M = c.sp->M, W = c.sp->W;
    // Default for np.linspace-like windows (the window of a position is arithmetic); other windows run
    // the delta formulation, which was as fast as a version of this kernel with a window search per
    // piece end (measured on C2: 0.97 vs 0.97 ms per step).  TSKB_SUM_VARIANT=runs also takes calls
    // with 2-5 columns (one pass per column); any other value of the variable selects another kernel.
    const char *variant = getenv("TSKB_SUM_VARIANT");
    const bool forced = variant != nullptr && variant[0] == 'r';
    if (variant != nullptr && !forced) return false;
    if (!c.sumP.skip_zero_bl || P.npp == 0 || P.T == 0 || W == 0) return false;
    // one pass over the pieces per column: with several columns the delta kernel (one pass) wins
    if (M >= COLS_KERNEL_MIN || (M > 1 && !forced)) return false;
    {
        // too many windows in the engine's range for enough copies of the bins in L2: the deltas scale
        // better (decided before the per-edge exactness check below, which is a host loop over W)
        const double *w = c.sp->windows;
        const size_t i0 = std::upper_bound(w, w + W + 1, P.range_left) - w;
        const size_t i1 = std::lower_bound(w, w + W + 1, P.range_right) - w;
        if (i1 > i0 + RUNS_MAX_WL) return false;
    }
    bool exact = false;
    RunArgs b = make_run_args(c.sp->windows, W, c.d_windows, exact, P.range_left, P.range_right);
    if (!exact) return false;  // the kernel is arithmetic on uniform edges only
    if (b.Wl > RUNS_MAX_WL + 3) return false;
    Arena &A = P.arena;
    ensure_piece_positions(P, c.s);
    // copies: reductions that meet at one address serialise in L2.  C2, summary + reduce per step
    // (profiles/r2x): 32 copies 0.96 ms, 128 0.89, 512 0.87, 1024 0.87
    uint32_t copies = 1024;
    if (const char *e = getenv("TSKB_RUN_COPIES")) copies = (uint32_t) std::max(1, atoi(e));  // experiments
    uint32_t pow2 = 1;
    while (pow2 * 2 <= copies) pow2 *= 2;
    copies = pow2;
    const uint32_t block = 2 * b.Wl + 1;
    while (copies > 1 && (size_t) copies * block * sizeof(double) > (size_t(16) << 20)) copies >>= 1;
    double *bins = A.get<double>((size_t) copies * block);
    double *RC = A.get<double>(block);
    b.bins = bins; b.ncols = 1; b.copy_mask = copies - 1;
    launch_sweep<V>(c, pval);
    TSKB_CK(cudaEventRecord(P.ev[2], c.s));
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, P.device);
    auto kern = k_branch_summary_runs<STAT, V>;
    const size_t smem = 0;
    TSKB_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RUN_TB, smem));
    int mult = std::max(per_sm, 1);
    if (const char *e = getenv("TSKB_RUN_GRID_MULT")) mult = std::max(1, atoi(e));  // experiments
    const uint32_t nblocks = (P.npp / RUN_IPT + RUN_TB - 1) / RUN_TB;
    for (uint32_t m = 0; m < M; m++) {
        TSKB_CK(cudaMemsetAsync(bins, 0, (size_t) copies * block * sizeof(double), c.s));
        kern<<<std::min<uint32_t>(nblocks, (uint32_t) (sms * mult)), RUN_TB, smem, c.s>>>(P.npp, P.q_x0.p, P.q_x1.p,
            P.q_bl.p, pval, c.sumP, totals, c.sumP.cols, m, b);
        TSKB_CK_LAUNCH();
        c.launches++;
        if (m == 0) TSKB_CK(cudaEventRecord(P.ev[3], c.s));
        k_runs_reduce<<<(block + 31) / 32, 1024, 0, c.s>>>(bins, block, copies, RC);
        k_runs_finalize<<<1, 1024, 0, c.s>>>(RC, c.d_windows, W, b.wlo, b.Wl, 1, m, M,
            (c.sp->options & TSKB_STAT_SPAN_NORMALISE) ? 1 : 0, c.d_result);
        TSKB_CK_LAUNCH();
        c.launches += 2;
    }
    TSKB_CK(cudaEventRecord(P.ev[4], c.s));
    return true;
}

template <int STAT, class V>
void run_branch(CallCtx &c, V *pval, V totals) {
    if (run_branch_runs<STAT, V>(c, pval, totals)) return;
    if (run_branch_bins<STAT, V>(c, pval, totals)) return;
    const Plan &P = *c.P;
    const uint32_t M = c.sp->M;
    Arena &A = P.arena;
    const uint32_t Tp1 = P.T + 1;
    const size_t col_bytes = (size_t) Tp1 * sizeof(double);
    // 6 or more columns: lanes-are-columns kernel, at most 32 columns per pass
    uint32_t cols_min = COLS_KERNEL_MIN;
    if (const char *e = getenv("TSKB_COLS_MIN")) cols_min = (uint32_t) std::max(1, atoi(e));  // experiments
    const bool by_cols = M >= cols_min && getenv("TSKB_NO_COLS_KERNEL") == nullptr;
    const char *cols_variant = getenv("TSKB_COLS_VARIANT");  // experiments: "old" = processing-order walk
    const bool by_pos = by_cols && !(cols_variant != nullptr && cols_variant[0] == 'o');
    // window runs instead of per-breakpoint deltas: finite summaries, bins of a 32-column pass within 64 MB
    const bool pos_runs = by_pos && c.sumP.skip_zero_bl && P.T > 0 && c.sp->W > 0
                          && (2 * (size_t) c.sp->W + 1) * std::min<uint32_t>(M, 32) * sizeof(double) <= (size_t(64) << 20)
                          && !(cols_variant != nullptr && cols_variant[0] == 'd');
    size_t budget = DELTA_BUDGET;
    if (by_pos && !pos_runs) {
        // the deltas of a column group may be large (their hot front stays in L2): up to a third of
        // the free memory, so that all columns go through in as few walks over the pieces as possible.
        // (cudaMemGetInfo costs milliseconds in a process that holds tens of GB and peer mappings: only here.)
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            budget = std::max(budget, std::min<size_t>((free_b + A.cap) / 3, size_t(24) << 30));
        }
    }
    uint32_t mc = (uint32_t) std::min<size_t>(M, std::max<size_t>(1, budget / col_bytes));
    if (by_cols) mc = std::min<uint32_t>(mc, 32);
    if (pos_runs) {
        ensure_summary_order(P, c.s);
        const uint32_t W = c.sp->W;
        bool exact = false;
        RunArgs b = make_run_args(c.sp->windows, W, c.d_windows, exact, P.range_left, P.range_right);
        const uint32_t ncmax = std::min<uint32_t>(M, 32);
        const size_t block_max = (2 * (size_t) b.Wl + 1) * ncmax;
        uint32_t copies = 64;
        if (const char *e = getenv("TSKB_RUN_COPIES")) copies = (uint32_t) std::max(1, atoi(e));  // experiments
        while (copies > 1 && copies * block_max * sizeof(double) > (size_t(64) << 20)) copies >>= 1;
        uint32_t pow2 = 1;
        while (pow2 * 2 <= copies) pow2 *= 2;
        copies = pow2;
        double *bins = A.get<double>(copies * block_max);
        double *RC = A.get<double>(block_max);
        b.bins = bins; b.copy_mask = copies - 1;
        const uint32_t nchunks = (P.nsp + BYPOS_CHUNK - 1) / BYPOS_CHUNK;
        // concurrent warps far apart along the genome: a multiplier coprime to the number of chunks
        uint32_t mul = std::max<uint32_t>(1, (uint32_t) (nchunks * 0.6180339887) | 1u);
        auto gcd = [](uint32_t x, uint32_t y) { while (y) { const uint32_t t = x % y; x = y; y = t; } return x; };
        while (nchunks > 1 && gcd(mul, nchunks) != 1) mul += 2;
        if (getenv("TSKB_RUN_SEQUENTIAL") != nullptr || nchunks <= 1) mul = 1;  // experiments
        b.chunk_mul = mul;
        launch_sweep<V>(c, pval);
        TSKB_CK(cudaEventRecord(P.ev[2], c.s));
        for (uint32_t m0 = 0; m0 < M; m0 += 32) {
            const uint32_t nc = std::min(M, m0 + 32) - m0;
            const uint32_t block = (2 * b.Wl + 1) * nc;
            b.ncols = nc;
            TSKB_CK(cudaMemsetAsync(bins, 0, (size_t) copies * block * sizeof(double), c.s));
            if (P.nsp > 0) {
                uint32_t CP = 1;
                while (CP < nc) CP <<= 1;  // lanes per piece: the columns rounded up to a power of two
                k_branch_summary_bypos_runs<STAT, V><<<grid_for((size_t) nchunks * 32, TB), TB, 0, c.s>>>(P.nsp,
                    P.so_slot.p, P.so_bp0.p, P.so_bp1.p, P.so_bl.p, P.bp_pos.p, pval, c.sumP, totals, c.sumP.cols, m0,
                    nc, CP, b, exact ? 1 : 0, nchunks);
                TSKB_CK_LAUNCH();
                c.launches++;
            }
            if (m0 == 0) TSKB_CK(cudaEventRecord(P.ev[3], c.s));
            k_runs_reduce<<<(block + 31) / 32, 1024, 0, c.s>>>(bins, block, copies, RC);
            k_runs_finalize<<<nc, 1024, 0, c.s>>>(RC, c.d_windows, W, b.wlo, b.Wl, nc, m0, M,
                (c.sp->options & TSKB_STAT_SPAN_NORMALISE) ? 1 : 0, c.d_result);
            TSKB_CK_LAUNCH();
            c.launches += 2;
        }
        TSKB_CK(cudaEventRecord(P.ev[4], c.s));
        return;
    }
    if (by_pos) {
        ensure_summary_order(P, c.s);
        double *Dp = A.get<double>((size_t) mc * Tp1);  // [breakpoint][column] deltas
        const uint32_t T = Tp1 - 1;
        const uint32_t nchunks = (T + COLSCAN_ROWS - 1) / COLSCAN_ROWS;
        double *partial = A.get<double>((size_t) std::max<uint32_t>(nchunks, 1) * mc);
        launch_sweep<V>(c, pval);
        TSKB_CK(cudaEventRecord(P.ev[2], c.s));
        TSKB_CK(cudaMemsetAsync(c.d_result, 0, (size_t) c.sp->W * M * sizeof(double), c.s));
        for (uint32_t m0 = 0; m0 < M; m0 += mc) {
            const uint32_t nc = std::min(M, m0 + mc) - m0;
            TSKB_CK(cudaMemsetAsync(Dp, 0, (size_t) nc * col_bytes, c.s));
            DeltaOut ox = { Dp, Tp1, c.sumP.cols };
            if (P.nsp > 0) {
                const uint32_t nwarps = (P.nsp + BYPOS_CHUNK - 1) / BYPOS_CHUNK;
                uint32_t CP = 1;
                while (CP < nc) CP <<= 1;  // lanes per piece: the columns rounded up to a power of two
                k_branch_summary_bypos<STAT, V><<<grid_for((size_t) nwarps * 32, TB), TB, 0, c.s>>>(P.nsp,
                    P.so_slot.p, P.so_bp0.p, P.so_bp1.p, P.so_bl.p, pval, c.sumP, totals, ox, m0, nc, CP);
                TSKB_CK_LAUNCH();
                c.launches++;
            }
            if (m0 == 0) TSKB_CK(cudaEventRecord(P.ev[3], c.s));
            if (T > 0) {
                const int g = grid_for((size_t) nchunks * 32, TB);
                k_colscan_partial<<<g, TB, 0, c.s>>>(Dp, T, nc, partial);
                k_colscan_offsets<<<1, 1024, 0, c.s>>>(partial, nchunks, nc);
                k_colscan_integrate<<<g, TB, 0, c.s>>>(Dp, T, nc, partial, P.bp_pos.p, c.d_windows, c.sp->W,
                    m0, M, c.d_result);
                TSKB_CK_LAUNCH();
                c.launches += 3;
            }
            if (c.sp->options & TSKB_STAT_SPAN_NORMALISE) {
                k_span_divide<<<grid_for((size_t) c.sp->W * nc, TB), TB, 0, c.s>>>(c.d_windows, c.sp->W, M, m0, nc,
                    c.d_result);
                TSKB_CK_LAUNCH();
                c.launches++;
            }
        }
        TSKB_CK(cudaEventRecord(P.ev[4], c.s));
        return;
    }
    double *D = A.get<double>((size_t) mc * Tp1);
    double *Dx = by_cols ? A.get<double>((size_t) mc * Tp1) : nullptr;  // [breakpoint][column] deltas
    size_t scan_bytes = 0;
    TSKB_CK(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, D, D, (int) std::max<uint32_t>(Tp1 - 1, 1), c.s));
    void *scan_tmp = A.get<char>(scan_bytes);
    DeltaOut out = { D, Tp1, c.sumP.cols };
    launch_sweep<V>(c, pval);
    TSKB_CK(cudaEventRecord(P.ev[2], c.s));
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, P.device);
    TSKB_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_branch_summary<STAT, V>, SUM_TB, 0));
    const uint32_t ntiles = (P.npp + SUM_TILE - 1) / SUM_TILE;
    for (uint32_t m0 = 0; m0 < M; m0 += mc) {
        const uint32_t m1 = std::min(M, m0 + mc);
        if (by_cols) {
            const uint32_t nc = m1 - m0;
            TSKB_CK(cudaMemsetAsync(Dx, 0, (size_t) nc * col_bytes, c.s));
            DeltaOut ox = { Dx, Tp1, c.sumP.cols };
            int per_sm_c = 1;
            TSKB_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_c, k_branch_summary_cols<STAT, V>, TB, 0));
            if (P.npp > 0) {
                const uint32_t nchunks = (P.npp + 31) / 32;
                const uint32_t grid = std::min<uint32_t>((nchunks + 7) / 8, (uint32_t) (sms * std::max(per_sm_c, 1)));
                k_branch_summary_cols<STAT, V><<<grid, TB, 0, c.s>>>(P.npp, P.q_bp0.p, P.q_bp1.p, P.q_bl.p,
                    pval, c.sumP, totals, ox, m0, nc);
                TSKB_CK_LAUNCH();
                c.launches++;
            }
            k_transpose_deltas<<<(Tp1 + 31) / 32, dim3(32, 8), 0, c.s>>>(Dx, Tp1, nc, D);
            TSKB_CK_LAUNCH();
            c.launches++;
            if (m0 == 0) TSKB_CK(cudaEventRecord(P.ev[3], c.s));
            finish_columns(c, D, Tp1, m0, nc, scan_tmp, scan_bytes);
            continue;
        }
        TSKB_CK(cudaMemsetAsync(D, 0, (size_t) (m1 - m0) * col_bytes, c.s));
        if (ntiles > 0) {
            // many more CTAs than are resident: later ones start as earlier ones finish (measured 8 % faster
            // than exactly-resident persistent CTAs)
            int mult = std::max(per_sm, 16);
            if (const char *e = getenv("TSKB_SUM_GRID_MULT")) mult = std::max(1, atoi(e));  // experiments
            // default: one piece per lane (k_branch_summary).  TSKB_SUM_VARIANT=c4 selects the
            // thread-contiguous kernel: measured on C2 (profiles/r2a_ab_summary.txt) it removes the MIO-throttle
            // stalls (19.6 -> 1.1 cycles per issue) and a third of the instructions, and is 5 % SLOWER
            // (1.014 vs 0.967 ms per step): the kernel is bound by the rate of scattered fp64 reductions in
            // L2 (half of them cross the die-to-die fabric), not by instruction issue.
            const char *variant = getenv("TSKB_SUM_VARIANT");
            if (variant == nullptr || variant[0] != 'c') {
                k_branch_summary<STAT, V><<<std::min<uint32_t>(ntiles, (uint32_t) (sms * mult)), SUM_TB, 0, c.s>>>(
                    P.npp, P.q_bp0.p, P.q_bp1.p, P.q_bl.p, pval, c.sumP, totals, out, m0, m1);
            } else {
                const uint32_t nblocks = (P.npp / SUMC_IPT + SUMC_TB - 1) / SUMC_TB;
                k_branch_summary_c4<STAT, V><<<std::min<uint32_t>(nblocks, (uint32_t) (sms * mult)), SUMC_TB, 0, c.s>>>(
                    P.npp, P.q_bp0.p, P.q_bp1.p, P.q_bl.p, pval, c.sumP, totals, out, m0, m1);
            }
            TSKB_CK_LAUNCH();
            c.launches++;
        }
        if (m0 == 0) TSKB_CK(cudaEventRecord(P.ev[3], c.s));
        finish_columns(c, D, Tp1, m0, m1 - m0, scan_tmp, scan_bytes);
    }
    TSKB_CK(cudaEventRecord(P.ev[4], c.s));
}

template <int STAT, class V>
void run_site(CallCtx &c, V *pval, V totals) {
    const Plan &P = *c.P;
    const uint32_t W = c.sp->W, M = c.sp->M;
    Arena &A = P.arena;
    launch_sweep<V>(c, pval);
    TSKB_CK(cudaEventRecord(P.ev[2], c.s));
    const uint32_t nsites = P.site_hi - P.site_lo;
    const uint32_t nsplit = std::max<uint32_t>(1, (592 + W - 1) / W);
    double *partial = A.get<double>((size_t) W * nsplit * M);
    double *R = A.get<double>((size_t) M * std::max<uint32_t>(nsites, 1));
    V *scratch = A.get<V>(P.total_alleles + 1);
    if (nsites) {
        k_site_summary<STAT, V><<<grid_for(nsites, 128), 128, 0, c.s>>>(P.site_lo, nsites,
            P.site_moff.p, P.site_aoff.p, P.mut_src.p, P.mut_allele.p, P.mut_alt.p, pval, totals,
            scratch, c.sumP, R);
        TSKB_CK_LAUNCH();
        c.launches++;
    }
    TSKB_CK(cudaEventRecord(P.ev[3], c.s));
    k_window_site<<<dim3(W, nsplit), TB, 0, c.s>>>(c.d_windows, nsplit, P.site_pos.p, P.site_lo,
        nsites, R, M, partial);
    TSKB_CK_LAUNCH();
    k_window_final<<<grid_for((size_t) W * M, TB), TB, 0, c.s>>>(partial, c.d_windows, W, nsplit, M,
        (c.sp->options & TSKB_STAT_SPAN_NORMALISE) ? 1 : 0, c.d_result);
    TSKB_CK_LAUNCH();
    c.launches += 2;
    TSKB_CK(cudaEventRecord(P.ev[4], c.s));
}

template <int STAT, class V>
void run_node(CallCtx &c, V *pval, V totals) {
    const Plan &P = *c.P;
    const uint32_t W = c.sp->W, M = c.sp->M, N = (uint32_t) P.N;
    launch_sweep<V>(c, pval);
    TSKB_CK(cudaEventRecord(P.ev[2], c.s));
    const int span = (c.sp->options & TSKB_STAT_SPAN_NORMALISE) ? 1 : 0;
    TSKB_CK(cudaMemsetAsync(c.d_result, 0, (size_t) W * N * M * sizeof(double), c.s));
    if (P.npp) {
        k_node_pieces<STAT, V><<<grid_for(P.npp, TB), TB, 0, c.s>>>(P.npp, P.q_node.p, P.q_bp0.p, P.q_bp1.p,
            P.bp_pos.p, pval, c.sumP, totals, c.d_windows, W, N, span, c.d_result);
        TSKB_CK_LAUNCH();
        c.launches++;
    }
    TSKB_CK(cudaEventRecord(P.ev[3], c.s));
    if (N) {
        k_node_initial<STAT, V><<<grid_for(N, TB), TB, 0, c.s>>>(N, P.node_first_bp.p, P.d_sample_index.p,
            pval + P.npp, P.num_samples, P.bp_pos.p, P.range_left, c.sumP, totals, c.d_windows, W, span,
            c.d_result);
        TSKB_CK_LAUNCH();
        c.launches++;
    }
    TSKB_CK(cudaEventRecord(P.ev[4], c.s));
}

// the spectrum's shape on the device (packed coordinates: the sets' dimensions)
inline AfsShape afs_shape(CallCtx &c) {
    AfsShape sh = {};
    const StatSpec &sp = *c.sp;
    if (sp.afs_dims != nullptr) {
        uint32_t *d = c.P->arena.get<uint32_t>(sp.afs_nsets);
        TSKB_CK(cudaMemcpyAsync(d, sp.afs_dims, sp.afs_nsets * sizeof(uint32_t), cudaMemcpyHostToDevice, c.s));
        sh.packed = 1;
        sh.nsets = sp.afs_nsets;
        sh.dims = d;
    }
    return sh;
}

// joint allele frequency spectrum, site mode
template <class V>
void run_afs_site(CallCtx &c, V *pval, V totals) {
    const Plan &P = *c.P;
    const uint32_t W = c.sp->W;
    const size_t afs_size = c.sp->afs_size;
    Arena &A = P.arena;
    const AfsShape sh = afs_shape(c);
    launch_sweep<V>(c, pval);
    TSKB_CK(cudaEventRecord(P.ev[2], c.s));
    const uint32_t nsites = P.site_hi - P.site_lo;
    V *scratch = A.get<V>(P.total_alleles + 1);
    TSKB_CK(cudaMemsetAsync(c.d_result, 0, (size_t) W * afs_size * sizeof(double), c.s));
    if (nsites) {
        k_site_afs<V><<<grid_for(nsites, 128), 128, 0, c.s>>>(P.site_lo, nsites, P.site_moff.p, P.site_aoff.p,
            P.mut_src.p, P.mut_allele.p, P.mut_alt.p, pval, totals, scratch, c.sumP, sh, P.num_samples,
            P.site_pos.p, c.d_windows, W, afs_size, c.d_result);
        TSKB_CK_LAUNCH();
        c.launches++;
    }
    TSKB_CK(cudaEventRecord(P.ev[3], c.s));
    if (c.sp->options & TSKB_STAT_SPAN_NORMALISE) {
        k_afs_span_normalise<<<grid_for((size_t) W * afs_size, TB), TB, 0, c.s>>>(c.d_windows, W, afs_size,
            c.d_result);
        TSKB_CK_LAUNCH();
        c.launches++;
    }
    TSKB_CK(cudaEventRecord(P.ev[4], c.s));
}

template <class V>
void run_afs_branch(CallCtx &c, V *pval) {
    const Plan &P = *c.P;
    const uint32_t W = c.sp->W, NTW = c.sp->num_time_windows;
    const size_t afs_size = c.sp->afs_size;
    const AfsShape sh = afs_shape(c);
    const double *d_tw = nullptr;
    const int32_t *q_node = nullptr;
    if (c.sp->time_windows != nullptr) {  // the caller made sure this plan keeps the node of every piece
        double *t = P.arena.get<double>(NTW + 1);
        TSKB_CK(cudaMemcpyAsync(t, c.sp->time_windows, (NTW + 1) * sizeof(double), cudaMemcpyHostToDevice, c.s));
        d_tw = t;
        q_node = P.q_node.p;
    }
    launch_sweep<V>(c, pval);
    TSKB_CK(cudaEventRecord(P.ev[2], c.s));
    TSKB_CK(cudaMemsetAsync(c.d_result, 0, (size_t) W * NTW * afs_size * sizeof(double), c.s));
    if (P.npp) {
        k_branch_afs<V><<<grid_for(P.npp, TB), TB, 0, c.s>>>(P.npp, P.q_bp0.p, P.q_bp1.p, P.q_eff.p, P.q_bl.p,
            P.bp_pos.p, pval, c.sumP, sh, P.num_samples, P.range_left, c.d_windows, W, afs_size, q_node, P.time.p,
            d_tw, NTW, c.d_result);
        TSKB_CK_LAUNCH();
        c.launches++;
    }
    TSKB_CK(cudaEventRecord(P.ev[3], c.s));
    if (c.sp->options & TSKB_STAT_SPAN_NORMALISE) {
        k_afs_span_normalise<<<grid_for((size_t) W * NTW * afs_size, TB), TB, 0, c.s>>>(c.d_windows, W,
            NTW * afs_size, c.d_result);
        TSKB_CK_LAUNCH();
        c.launches++;
    }
    TSKB_CK(cudaEventRecord(P.ev[4], c.s));
}

// relatedness vector: sweep, then one push launch per height and window, tallest height first
template <class V>
void run_relvec(CallCtx &c, V *pval) {
    if constexpr (std::is_same<typename V::scalar, double>::value) {
        const Plan &P = *c.P;
        const StatSpec &sp = *c.sp;
        const uint32_t W = sp.W, K = sp.K, nf = (uint32_t) sp.num_focal;
        Arena &A = P.arena;
        launch_sweep<V>(c, pval);
        TSKB_CK(cudaEventRecord(P.ev[2], c.s));
        const size_t slots = (size_t) P.npp + P.num_samples + 1;
        const size_t gn_stride = sp.focal_needs_nodes ? (size_t) P.N * K : 0;
        // windows in chunks whose accumulators fit 8 GB; one push launch per height and chunk
        const size_t per_window = slots * sizeof(V) + gn_stride * sizeof(double);
        uint32_t wc_max = (uint32_t) std::max<size_t>(1, (size_t(8) << 30) / per_window);
        if (wc_max > W) wc_max = W;
        V *G = A.get<V>(slots * wc_max);
        int32_t *d_focal = A.get<int32_t>(nf + 1);
        TSKB_CK(cudaMemcpyAsync(d_focal, sp.focal, (size_t) nf * sizeof(int32_t), cudaMemcpyHostToDevice, c.s));
        double *Gn = gn_stride ? A.get<double>(gn_stride * wc_max) : nullptr;
        for (uint32_t w0 = 0; w0 < W; w0 += wc_max) {
            const uint32_t wc = W - w0 < wc_max ? W - w0 : wc_max;
            TSKB_CK(cudaMemsetAsync(G, 0, slots * wc * sizeof(V), c.s));
            if (Gn) TSKB_CK(cudaMemsetAsync(Gn, 0, gn_stride * wc * sizeof(double), c.s));
            for (uint32_t h = P.nheights; h-- > 0;) {
                const uint32_t lo = P.level_begin[h], cnt = P.level_begin[h + 1] - lo;
                if (cnt == 0) continue;
                k_relvec_push<V><<<grid_for(cnt, TB), TB, 0, c.s>>>(lo, cnt, P.q_bp0.p, P.q_bp1.p, P.q_bl.p,
                    P.bp_pos.p, P.q_off.p, P.refs.p, pval, G, slots, c.d_windows, W, w0, wc, P.q_node.p, Gn,
                    gn_stride, K);
                TSKB_CK_LAUNCH();
                c.launches++;
            }
            if ((size_t) nf * K) {
                k_relvec_out<V><<<grid_for((size_t) nf * K * wc, TB), TB, 0, c.s>>>(d_focal, nf, K, wc,
                    P.d_sample_index.p, G, slots, P.npp, Gn, gn_stride, c.d_result + (size_t) w0 * nf * K);
                TSKB_CK_LAUNCH();
                c.launches++;
            }
        }
        TSKB_CK(cudaEventRecord(P.ev[3], c.s));
        TSKB_CK(cudaEventRecord(P.ev[4], c.s));
    }
}

template <int STAT, class V>
void run_phases(CallCtx &c, V *pval, V totals) {
    if (c.sp->options & TSKB_STAT_NODE) {
        run_node<STAT, V>(c, pval, totals);
    } else if (c.sp->options & TSKB_STAT_BRANCH) {
        run_branch<STAT, V>(c, pval, totals);
    } else {
        run_site<STAT, V>(c, pval, totals);
    }
}

template <class V>
int run_impl(const Plan &P, const StatSpec &sp) {
    cudaStream_t s = P.stream;
    const uint32_t K = sp.K, M = sp.M, W = sp.W;
    const bool timing = getenv("TSKB_TIMING") != nullptr;  // host-side phases of the call, to stderr
    const auto t_enter = std::chrono::steady_clock::now();
    Arena &A = P.arena;
    A.reset();
    CallCtx c = {};
    c.P = &P;
    c.sp = &sp;
    c.s = s;
    TSKB_CK(cudaEventRecord(P.ev[0], s));

    // ---- phase 0: weights.  State slots: [npp pieces | one INIT slot per sample | zero slot]
    constexpr bool WEIGHTED = std::is_same<typename V::scalar, double>::value;
    uint64_t total = 0;
    std::vector<uint32_t> h_off(K + 1, 0);
    for (uint32_t k = 0; k < K && !WEIGHTED; k++) {
        total += sp.sizes[k];
        h_off[k + 1] = (uint32_t) total;
    }
    V *pval = A.get<V>((size_t) P.npp + P.num_samples + 1);
    V *init = pval + P.npp;
    const int32_t *d_sets = sp.sets;
    if (!sp.sets_on_device && !WEIGHTED) {
        int32_t *tmp_sets = A.get<int32_t>(total);
        TSKB_CK(cudaMemcpyAsync(tmp_sets, sp.sets, total * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        d_sets = tmp_sets;
    }

    SumP &sumP = c.sumP;
    sumP.K = (int) K;
    sumP.M = (int) M;
    sumP.polarised = (sp.options & TSKB_STAT_POLARISED) ? 1 : 0;
    sumP.skip_zero_bl = 1;
    sumP.timing_hack = getenv("TSKB_SUM_HACK") != nullptr ? atoi(getenv("TSKB_SUM_HACK")) : 0;
    V totals;
    for (int k = 0; k < V::N; k++) totals.v[k] = 0;
    for (uint32_t k = 0; k < K; k++) {
        if constexpr (WEIGHTED) {
            sumP.n[k] = (double) P.num_samples;
            totals.v[k] = sp.column_totals[k];  // total_weight of general_stat (trees.c:1406-1415)
        } else {
            sumP.n[k] = (double) sp.sizes[k];
            totals.v[k] = (int32_t) sp.sizes[k];
        }
    }
    // All small per-call inputs travel in ONE host-to-device copy:
    //   [0] validation key (all ones)  [8] duplicate flag, sweep error flag  [16] completion counters
    //   [24] set offsets (K + 1, padded)  | result columns (M)  | window edges (W + 1)
    const uint32_t Mc = (sp.stat_id == STAT_AFS || sp.stat_id == STAT_REL_VECTOR) ? 0 : M;  // the spectrum has no per-column parameters
    const size_t off_bytes = ((size_t) (K + 1) * sizeof(uint32_t) + 7) & ~size_t(7);
    const size_t o_off = 24, o_cols = o_off + off_bytes, o_win = o_cols + (size_t) Mc * sizeof(ColP);
    const size_t stage_bytes = o_win + (size_t) (W + 1) * sizeof(double);
    std::vector<unsigned long long> stage((stage_bytes + 7) / 8, 0);
    char *hs = reinterpret_cast<char *>(stage.data());
    stage[0] = ~0ull;
    memcpy(hs + o_off, h_off.data(), (K + 1) * sizeof(uint32_t));
    ColP *cols = reinterpret_cast<ColP *>(hs + o_cols);
    for (uint32_t m = 0; m < Mc; m++) {
        ColP &q = cols[m];
        int32_t t[4] = { (int32_t) (m < K ? m : 0), 0, 0, 0 };
        for (uint32_t a = 0; a < sp.tuple; a++) t[a] = sp.indexes[(size_t) m * sp.tuple + a];
        if (sp.stat_id == STAT_TABULATED) t[0] = 0;
        q.i = t[0]; q.j = t[1]; q.k = t[2]; q.l = t[3];
        if constexpr (WEIGHTED) {
            const double n = (double) P.num_samples;
            q.ni = q.nj = q.nk = q.nl = 0.0;
            q.inv = 1.0;
            if (sp.stat_id == STAT_TRAIT_COV) {
                q.ni = 2 * (n - 1) * (n - 1);  // trees.c:3972
            } else if (sp.stat_id == STAT_REL_WEIGHTED || sp.stat_id == STAT_REL_WEIGHTED_NC
                       || sp.stat_id == STAT_REL_SIDE) {
                q.ni = sp.column_totals[t[0]];
                q.nj = sp.column_totals[t[1]];
            }
            continue;
        }
        q.ni = (double) sp.sizes[t[0]]; q.nj = (double) sp.sizes[t[1]];
        q.nk = (double) sp.sizes[t[2]]; q.nl = (double) sp.sizes[t[3]];
        q.inv = 1.0 / column_denominator(sp.stat_id, q);
        if (!std::isfinite(q.inv)) sumP.skip_zero_bl = 0;
    }
    memcpy(hs + o_win, sp.windows, (W + 1) * sizeof(double));
    char *ds = A.get<char>(stage_bytes);
    TSKB_CK(cudaMemcpyAsync(ds, hs, stage_bytes, cudaMemcpyHostToDevice, s));
    unsigned long long *d_verr = reinterpret_cast<unsigned long long *>(ds);
    int *d_dup = reinterpret_cast<int *>(ds + 8);
    c.d_err = d_dup + 1;
    c.d_counters = reinterpret_cast<uint32_t *>(ds + 16);
    uint32_t *d_off = reinterpret_cast<uint32_t *>(ds + o_off);
    sumP.cols = reinterpret_cast<const ColP *>(ds + o_cols);
    c.d_windows = reinterpret_cast<double *>(ds + o_win);
    TSKB_CK(cudaMemsetAsync(init, 0, ((size_t) P.num_samples + 1) * sizeof(V), s));
    if constexpr (WEIGHTED) {
        (void) d_sets; (void) d_off; (void) d_verr; (void) total;
        double *d_w = A.get<double>((size_t) P.num_samples * K);
        TSKB_CK(cudaMemcpyAsync(d_w, sp.weights, (size_t) P.num_samples * K * sizeof(double),
            cudaMemcpyHostToDevice, s));
        if (P.num_samples) {
            k_init_weights<V><<<grid_for(P.num_samples, TB), TB, 0, s>>>(d_w, P.num_samples, K, init);
            TSKB_CK_LAUNCH();
            c.launches++;
        }
    } else if (total) {
        k_set_weights<V><<<grid_for(total, TB), TB, 0, s>>>(d_sets, d_off, K, (uint32_t) total,
            P.d_sample_index.p, (int32_t) P.N, init, d_verr, d_dup);
        TSKB_CK_LAUNCH();
        c.launches++;
    }
    if (sp.stat_id == STAT_TRAIT_LM && sp.table_rows > 0) {
        double *d_tab = A.get<double>(sp.table_rows * M);
        TSKB_CK(cudaMemcpyAsync(d_tab, sp.f_table, sp.table_rows * M * sizeof(double),
            cudaMemcpyHostToDevice, s));
        sumP.table = d_tab;
    }
    if (sp.stat_id == STAT_TRAIT_LM || sp.stat_id == STAT_REL_SIDE) sumP.table_rows = (uint32_t) sp.table_rows;
    if (sp.stat_id == STAT_TABULATED) {
        double *d_tab = A.get<double>(sp.table_rows * M);
        TSKB_CK(cudaMemcpyAsync(d_tab, sp.f_table, sp.table_rows * M * sizeof(double),
            cudaMemcpyHostToDevice, s));
        for (uint64_t i = 0; i < sp.table_rows * M; i++) {
            if (!std::isfinite(sp.f_table[i])) sumP.skip_zero_bl = 0;
        }
        // A summary that is NaN / inf at some state: the reference's running sum takes 0 x f = NaN from a
        // node WITHOUT a branch above it too (trees.c:1339-1350).  The default plan drops those pieces; the
        // plan that keeps every piece (TSKB_INIT_NODE_MODE) reproduces it -- lowlevel.py retries there.
        if (!sumP.skip_zero_bl && (sp.options & TSKB_STAT_BRANCH) && !P.all_pieces) return TSKB_ERR_UNSUPPORTED;
        sumP.table = d_tab;
        sumP.table_rows = (uint32_t) sp.table_rows;
    }
    // node mode: one row per node and window (trees.c:1788-1918)
    const size_t result_size = sp.stat_id == STAT_AFS
                                   ? (size_t) W * sp.num_time_windows * sp.afs_size
                                   : (size_t) W * M * ((sp.options & TSKB_STAT_NODE) ? (size_t) P.N : 1);
    c.d_result = sp.result_on_device ? sp.result : A.get<double>(result_size);
    TSKB_CK(cudaEventRecord(P.ev[1], s));

    // ---- phases 1-3: sweep (+ branch summary), summary, finalize
    if constexpr (WEIGHTED) {
        switch (sp.stat_id) {
            case STAT_TRAIT_COV: run_phases<STAT_TRAIT_COV, V>(c, pval, totals); break;
            case STAT_TRAIT_CORR: run_phases<STAT_TRAIT_CORR, V>(c, pval, totals); break;
            case STAT_REL_WEIGHTED: run_phases<STAT_REL_WEIGHTED, V>(c, pval, totals); break;
            case STAT_REL_WEIGHTED_NC: run_phases<STAT_REL_WEIGHTED_NC, V>(c, pval, totals); break;
            case STAT_REL_SIDE: run_phases<STAT_REL_SIDE, V>(c, pval, totals); break;
            case STAT_TRAIT_LM: run_phases<STAT_TRAIT_LM, V>(c, pval, totals); break;
            case STAT_REL_VECTOR: run_relvec<V>(c, pval); break;
            case STAT_AFS:  // packed spectrum coordinates (more than 7 sample sets)
                if (sp.options & TSKB_STAT_BRANCH) {
                    run_afs_branch<V>(c, pval);
                } else {
                    run_afs_site<V>(c, pval, totals);
                }
                break;
            default: return TSKB_ERR_BAD_PARAM_VALUE;
        }
    } else
    switch (sp.stat_id) {
        case STAT_DIVERSITY: run_phases<STAT_DIVERSITY, V>(c, pval, totals); break;
        case STAT_SEGSITES: run_phases<STAT_SEGSITES, V>(c, pval, totals); break;
        case STAT_Y1: run_phases<STAT_Y1, V>(c, pval, totals); break;
        case STAT_DIVERGENCE: run_phases<STAT_DIVERGENCE, V>(c, pval, totals); break;
        case STAT_Y2: run_phases<STAT_Y2, V>(c, pval, totals); break;
        case STAT_F2: run_phases<STAT_F2, V>(c, pval, totals); break;
        case STAT_RELATEDNESS: run_phases<STAT_RELATEDNESS, V>(c, pval, totals); break;
        case STAT_RELATEDNESS_NC: run_phases<STAT_RELATEDNESS_NC, V>(c, pval, totals); break;
        case STAT_Y3: run_phases<STAT_Y3, V>(c, pval, totals); break;
        case STAT_F3: run_phases<STAT_F3, V>(c, pval, totals); break;
        case STAT_F4: run_phases<STAT_F4, V>(c, pval, totals); break;
        case STAT_AFS:
            if (sp.options & TSKB_STAT_BRANCH) {
                run_afs_branch<V>(c, pval);
            } else {
                run_afs_site<V>(c, pval, totals);
            }
            break;
        case STAT_TABULATED:
            if constexpr (std::is_same<V, IVec<1>>::value) {
                run_phases<STAT_TABULATED, V>(c, pval, totals);
                break;
            }
            return TSKB_ERR_UNSUPPORTED;
        default: return TSKB_ERR_BAD_PARAM_VALUE;
    }
    TSKB_CK(cudaEventRecord(P.ev[5], s));
    unsigned long long h_flags[2] = { ~0ull, 0 };
    TSKB_CK(cudaMemcpyAsync(h_flags, d_verr, sizeof(h_flags), cudaMemcpyDeviceToHost, s));
    if (!sp.result_on_device) {
        TSKB_CK(cudaMemcpyAsync(sp.result, c.d_result, result_size * sizeof(double),
            cudaMemcpyDeviceToHost, s));
    }
    TSKB_CK(cudaEventRecord(P.ev[6], s));
    const auto t_enqueued = std::chrono::steady_clock::now();
    TSKB_CK(cudaStreamSynchronize(s));
    if (timing) {
        const auto t_done = std::chrono::steady_clock::now();
        fprintf(stderr, "tskb timing: stat %d enqueue %.3f ms, wait %.3f ms\n", sp.stat_id,
            std::chrono::duration<double, std::milli>(t_enqueued - t_enter).count(),
            std::chrono::duration<double, std::milli>(t_done - t_enqueued).count());
    }
    float ms = 0;
    for (int q = 0; q < 6; q++) {
        TSKB_CK(cudaEventElapsedTime(&ms, P.ev[q], P.ev[q + 1]));
        P.stats.last_kernel_ms[q] = ms;
    }
    TSKB_CK(cudaEventElapsedTime(&ms, P.ev[0], P.ev[6]));
    P.stats.last_call_ms = ms;
    P.stats.last_launches = c.launches;
    // sample-set errors found on the device: same codes and precedence as the host check
    const unsigned long long h_verr = h_flags[0];
    const int h_dup = (int) (h_flags[1] & 0xffffffffull), h_err = (int) (h_flags[1] >> 32);
    if (h_verr != ~0ull) return (h_verr & 1ull) ? TSKB_ERR_BAD_SAMPLES : TSKB_ERR_NODE_OUT_OF_BOUNDS;
    if (h_dup) return TSKB_ERR_DUPLICATE_SAMPLE;
    if (h_err) {
        last_error_string() = "sweep: wait for the lower heights timed out";
        return TSKB_ERR_CUDA;
    }
    return 0;
}

// ---------------------------------------------------------------- general_stat with a host callback
// tsk_treeseq_general_stat (trees.c:2035-2095) calls the summary function f at every node update.  Here
// the sweep runs first (fp64 states: sums of the weight rows), the DISTINCT state vectors are collected
// on the device (K stable radix sorts of a permutation, one per state column, then a head scan), f is
// called on the host once per distinct vector (branch mode: f(x) + f(total - x) unless polarised,
// trees.c:1944-1972), and the table of its values is read back by the ordinary summary kernels, whose
// "state" is then the index of the piece's vector.  Site mode does the same over the allele states.
__device__ __forceinline__ unsigned long long gs_ordered_bits(double x) {
    unsigned long long b = (unsigned long long) __double_as_longlong(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

template <class V>
__global__ void k_gs_keys(uint32_t n, const uint32_t *__restrict__ perm, const V *__restrict__ vals, int k,
    unsigned long long *keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = gs_ordered_bits((double) pick<V>(vals[perm[i]], k));
}

template <class V>
__global__ void k_gs_heads(uint32_t n, const uint32_t *__restrict__ perm, const V *__restrict__ vals,
    uint32_t *head) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool h = i == 0;
    if (!h) {
        const V a = vals[perm[i]], b = vals[perm[i - 1]];
#pragma unroll
        for (int k = 0; k < V::N; k++) h |= __double_as_longlong(a.v[k]) != __double_as_longlong(b.v[k]);
    }
    head[i] = h ? 1u : 0u;
}

template <class V>
__global__ void k_gs_scatter(uint32_t n, const uint32_t *__restrict__ perm, const V *__restrict__ vals,
    const uint32_t *__restrict__ head, const uint32_t *__restrict__ rank_incl, int K, int32_t *idx, double *distinct) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = rank_incl[i] - 1;
    idx[perm[i]] = (int32_t) r;
    if (head[i]) {
        const V a = vals[perm[i]];
        for (int k = 0; k < K; k++) distinct[(size_t) r * K + k] = pick<V>(a, k);
    }
}

__global__ void k_gs_valid(uint32_t npp, const uint32_t *__restrict__ q_bp1, uint8_t *flag) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < npp) flag[j] = q_bp1[j] != NO_PIECE;
}

// allele states of every site (the loop of k_site_summary, states kept)
template <class V>
__global__ void k_site_states(uint32_t site_lo, uint32_t nsites, const uint32_t *site_moff,
    const uint32_t *site_aoff, const int32_t *mut_src, const uint16_t *mut_allele, const uint16_t *mut_alt,
    const V *pval, V totals, V *scratch) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nsites) return;
    const uint32_t site = site_lo + t;
    const uint32_t a0 = site_aoff[site], na = site_aoff[site + 1] - a0;
    scratch[a0] = totals;  // allele 0 starts at total_weight (trees.c:1548)
    for (uint32_t al = 1; al < na; al++) scratch[a0 + al] = ivec_zero<V>();
    for (uint32_t m = site_moff[site]; m < site_moff[site + 1]; m++) {
        const V x = pval[mut_src[m]];
        scratch[a0 + mut_allele[m]] = scratch[a0 + mut_allele[m]] + x;
        scratch[a0 + mut_alt[m]] = scratch[a0 + mut_alt[m]] - x;
    }
}

__global__ void k_site_apply(uint32_t site_lo, uint32_t nsites, const uint32_t *site_aoff,
    const int32_t *__restrict__ idx, const double *__restrict__ table, uint32_t M, int polarised, double *R) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nsites) return;
    const uint32_t site = site_lo + t;
    const uint32_t a0 = site_aoff[site], na = site_aoff[site + 1] - a0;
    for (uint32_t m = 0; m < M; m++) {
        double acc = 0.0;
        for (uint32_t al = polarised ? 1 : 0; al < na; al++) acc += table[(size_t) idx[a0 + al] * M + m];
        R[(size_t) m * nsites + t] = acc;
    }
}

// items (state slots) -> index of their distinct vector; the distinct vectors on the host
template <class V>
uint32_t distinct_states(const Plan &P, cudaStream_t s, const V *vals, const uint32_t *items, uint32_t n,
    int K, int32_t *idx, std::vector<double> &h_distinct) {
    if (n == 0) return 0;
    DevArray<uint32_t> perm_a, perm_b, head, rank;
    DevArray<unsigned long long> key_a, key_b;
    DevArray<char> tmp;
    perm_a.alloc(n); perm_b.alloc(n); head.alloc(n); rank.alloc(n); key_a.alloc(n); key_b.alloc(n);
    TSKB_CK(cudaMemcpyAsync(perm_a.p, items, (size_t) n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    size_t bytes = 0, scan_bytes = 0;
    TSKB_CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, key_a.p, key_b.p, perm_a.p, perm_b.p, n, 0, 64, s));
    TSKB_CK(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, head.p, rank.p, n, s));
    tmp.alloc(std::max(bytes, scan_bytes));
    uint32_t *pa = perm_a.p, *pb = perm_b.p;
    for (int k = K - 1; k >= 0; k--) {  // least significant column first: stable sorts
        k_gs_keys<V><<<grid_for(n, TB), TB, 0, s>>>(n, pa, vals, k, key_a.p);
        TSKB_CK_LAUNCH();
        TSKB_CK(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, key_a.p, key_b.p, pa, pb, n, 0, 64, s));
        std::swap(pa, pb);
    }
    k_gs_heads<V><<<grid_for(n, TB), TB, 0, s>>>(n, pa, vals, head.p);
    TSKB_CK_LAUNCH();
    TSKB_CK(cub::DeviceScan::InclusiveSum(tmp.p, scan_bytes, head.p, rank.p, n, s));
    uint32_t D = 0;
    TSKB_CK(cudaMemcpyAsync(&D, rank.p + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    TSKB_CK(cudaStreamSynchronize(s));
    DevArray<double> d_distinct;
    d_distinct.alloc((size_t) D * K);
    k_gs_scatter<V><<<grid_for(n, TB), TB, 0, s>>>(n, pa, vals, head.p, rank.p, K, idx, d_distinct.p);
    TSKB_CK_LAUNCH();
    h_distinct.resize((size_t) D * K);
    TSKB_CK(cudaMemcpyAsync(h_distinct.data(), d_distinct.p, (size_t) D * K * sizeof(double), cudaMemcpyDeviceToHost, s));
    TSKB_CK(cudaStreamSynchronize(s));
    return D;
}

template <class V>
int run_general_impl(const Plan &P, const GeneralSpec &g) {
    cudaStream_t s = P.stream;
    const uint32_t K = g.K, M = g.M, W = g.W;
    const bool branch = (g.options & TSKB_STAT_BRANCH) != 0;
    const int polarised = (g.options & TSKB_STAT_POLARISED) ? 1 : 0;
    Arena &A = P.arena;
    A.reset();
    CallCtx c = {};
    StatSpec sp = {};
    sp.stat_id = STAT_TABULATED; sp.K = 1; sp.M = M; sp.W = W; sp.windows = g.windows; sp.options = g.options;
    c.P = &P; c.sp = &sp; c.s = s;
    TSKB_CK(cudaEventRecord(P.ev[0], s));
    // per-call inputs: [0] flags / counters, then the window edges
    const size_t o_win = 32, stage_bytes = o_win + (size_t) (W + 1) * sizeof(double);
    std::vector<unsigned long long> stage((stage_bytes + 7) / 8, 0);
    memcpy(reinterpret_cast<char *>(stage.data()) + o_win, g.windows, (W + 1) * sizeof(double));
    char *ds = A.get<char>(stage_bytes);
    TSKB_CK(cudaMemcpyAsync(ds, stage.data(), stage_bytes, cudaMemcpyHostToDevice, s));
    c.d_err = reinterpret_cast<int *>(ds + 8);
    c.d_counters = reinterpret_cast<uint32_t *>(ds + 16);
    c.d_windows = reinterpret_cast<double *>(ds + o_win);
    // states: the weight rows of the samples (trees.c:1406-1415)
    V *pval = A.get<V>((size_t) P.npp + P.num_samples + 1);
    V *init = pval + P.npp;
    TSKB_CK(cudaMemsetAsync(init, 0, ((size_t) P.num_samples + 1) * sizeof(V), s));
    double *d_w = A.get<double>((size_t) P.num_samples * K);
    TSKB_CK(cudaMemcpyAsync(d_w, g.weights, (size_t) P.num_samples * K * sizeof(double), cudaMemcpyHostToDevice, s));
    if (P.num_samples) {
        k_init_weights<V><<<grid_for(P.num_samples, TB), TB, 0, s>>>(d_w, P.num_samples, K, init);
        TSKB_CK_LAUNCH();
        c.launches++;
    }
    std::vector<double> total(K, 0.0);  // total_weight, summed over the samples in order (trees.c:2003-2010)
    for (uint64_t j = 0; j < P.num_samples; j++) {
        for (uint32_t k = 0; k < K; k++) total[k] += g.weights[j * K + k];
    }
    V totals = ivec_zero<V>();
    for (uint32_t k = 0; k < K; k++) totals.v[k] = total[k];
    TSKB_CK(cudaEventRecord(P.ev[1], s));
    launch_sweep<V>(c, pval);
    TSKB_CK(cudaEventRecord(P.ev[2], s));
    // the vectors f is evaluated at: branch mode the states of the pieces, site mode the allele states
    const uint32_t nsites = P.site_hi - P.site_lo;
    const V *vals = pval;
    uint32_t n_items = 0;
    DevArray<uint32_t> items;
    int32_t *idx = nullptr;
    if (branch) {
        DevArray<uint8_t> flag;
        DevArray<uint32_t> iota, cnt;
        flag.alloc(P.npp); iota.alloc(P.npp); cnt.alloc(1); items.alloc(P.npp);
        idx = reinterpret_cast<int32_t *>(A.get<IVec<1>>((size_t) P.npp + P.num_samples + 1));
        TSKB_CK(cudaMemsetAsync(idx, 0, ((size_t) P.npp + P.num_samples + 1) * sizeof(int32_t), s));
        if (P.npp) {
            k_gs_valid<<<grid_for(P.npp, TB), TB, 0, s>>>(P.npp, P.q_bp1.p, flag.p);
            TSKB_CK_LAUNCH();
            size_t bytes = 0;
            cub::CountingInputIterator<uint32_t> it(0);
            TSKB_CK(cub::DeviceSelect::Flagged(nullptr, bytes, it, flag.p, items.p, cnt.p, (int) P.npp, s));
            DevArray<char> tmp;
            tmp.alloc(bytes);
            TSKB_CK(cub::DeviceSelect::Flagged(tmp.p, bytes, it, flag.p, items.p, cnt.p, (int) P.npp, s));
            TSKB_CK(cudaMemcpyAsync(&n_items, cnt.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            TSKB_CK(cudaStreamSynchronize(s));
        }
    } else {
        V *scratch = A.get<V>(P.total_alleles + 1);
        idx = A.get<int32_t>(P.total_alleles + 1);
        // the allele slots of the sites inside this plan's range are contiguous
        uint32_t a_lo = 0, a_hi = 0;
        TSKB_CK(cudaMemcpyAsync(&a_lo, P.site_aoff.p + P.site_lo, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        TSKB_CK(cudaMemcpyAsync(&a_hi, P.site_aoff.p + P.site_hi, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        TSKB_CK(cudaStreamSynchronize(s));
        n_items = a_hi - a_lo;
        items.alloc(std::max<uint32_t>(n_items, 1));
        if (n_items) {
            k_site_states<V><<<grid_for(nsites, 128), 128, 0, s>>>(P.site_lo, nsites, P.site_moff.p, P.site_aoff.p,
                P.mut_src.p, P.mut_allele.p, P.mut_alt.p, pval, totals, scratch);
            TSKB_CK_LAUNCH();
            std::vector<uint32_t> h(n_items);
            for (uint32_t i = 0; i < n_items; i++) h[i] = a_lo + i;
            TSKB_CK(cudaMemcpyAsync(items.p, h.data(), (size_t) n_items * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
            TSKB_CK(cudaStreamSynchronize(s));
        }
        vals = scratch;
    }
    std::vector<double> h_distinct;
    const uint32_t D = distinct_states<V>(P, s, vals, items.p, n_items, (int) K, idx, h_distinct);
    // f on the host, once per distinct vector; a non-zero return aborts the call and is handed back
    // (trees.c:1396-1399, 1441)
    std::vector<double> table((size_t) std::max<uint32_t>(D, 1) * M, 0.0), tmp_out(M), other(K);
    bool finite = true;
    for (uint32_t d = 0; d < D; d++) {
        double *row = table.data() + (size_t) d * M;
        int ret = g.f(K, h_distinct.data() + (size_t) d * K, M, row, g.params);
        if (ret != 0) return ret;
        if (branch && !polarised) {
            for (uint32_t k = 0; k < K; k++) other[k] = total[k] - h_distinct[(size_t) d * K + k];
            ret = g.f(K, other.data(), M, tmp_out.data(), g.params);
            if (ret != 0) return ret;
            for (uint32_t m = 0; m < M; m++) row[m] += tmp_out[m];
        }
        for (uint32_t m = 0; m < M; m++) finite &= std::isfinite(row[m]);
    }
    if (!finite && branch && !P.all_pieces) return TSKB_ERR_UNSUPPORTED;  // see STAT_TABULATED in run_impl
    double *d_tab = A.get<double>(table.size());
    TSKB_CK(cudaMemcpyAsync(d_tab, table.data(), table.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    c.d_result = A.get<double>((size_t) W * M);
    SumP &sumP = c.sumP;
    sumP.K = 1; sumP.M = (int) M; sumP.polarised = 1;  // the table already holds f(x) + f(total - x)
    sumP.skip_zero_bl = finite ? 1 : 0;
    sumP.table = d_tab; sumP.table_rows = std::max<uint32_t>(D, 1);
    // result columns: only their number matters to a tabulated summary
    std::vector<ColP> cols(M);
    for (uint32_t m = 0; m < M; m++) {
        cols[m] = ColP{};
        cols[m].inv = 1.0;
    }
    ColP *d_cols = A.get<ColP>(M);
    TSKB_CK(cudaMemcpyAsync(d_cols, cols.data(), M * sizeof(ColP), cudaMemcpyHostToDevice, s));
    sumP.cols = d_cols;
    if (branch) {
        c.skip_sweep = true;
        IVec<1> zero = ivec_zero<IVec<1>>();
        run_branch<STAT_TABULATED, IVec<1>>(c, reinterpret_cast<IVec<1> *>(idx), zero);
    } else {
        const uint32_t nsplit = std::max<uint32_t>(1, (592 + W - 1) / W);
        double *partial = A.get<double>((size_t) W * nsplit * M);
        double *R = A.get<double>((size_t) M * std::max<uint32_t>(nsites, 1));
        if (nsites) {
            k_site_apply<<<grid_for(nsites, 128), 128, 0, s>>>(P.site_lo, nsites, P.site_aoff.p, idx, d_tab, M, polarised, R);
            TSKB_CK_LAUNCH();
        }
        TSKB_CK(cudaEventRecord(P.ev[3], s));
        k_window_site<<<dim3(W, nsplit), TB, 0, s>>>(c.d_windows, nsplit, P.site_pos.p, P.site_lo, nsites, R, M, partial);
        k_window_final<<<grid_for((size_t) W * M, TB), TB, 0, s>>>(partial, c.d_windows, W, nsplit, M,
            (g.options & TSKB_STAT_SPAN_NORMALISE) ? 1 : 0, c.d_result);
        TSKB_CK_LAUNCH();
        TSKB_CK(cudaEventRecord(P.ev[4], s));
    }
    TSKB_CK(cudaEventRecord(P.ev[5], s));
    TSKB_CK(cudaMemcpyAsync(g.result, c.d_result, (size_t) W * M * sizeof(double), cudaMemcpyDeviceToHost, s));
    TSKB_CK(cudaEventRecord(P.ev[6], s));
    TSKB_CK(cudaStreamSynchronize(s));
    P.stats.last_launches = c.launches;
    P.stats.last_kernel_ms[7] = (double) D;  // distinct state vectors = calls of f (x 2 when not polarised)
    return 0;
}

}  // namespace

int run_general_stat(const Plan *plan, const GeneralSpec &g) {
    std::lock_guard<std::mutex> lock(plan->mu);
    TSKB_CK(cudaSetDevice(plan->device));
    if (g.K <= 1) return run_general_impl<DVec<1>>(*plan, g);
    if (g.K <= 2) return run_general_impl<DVec<2>>(*plan, g);
    if (g.K <= 4) return run_general_impl<DVec<4>>(*plan, g);
    if (g.K <= 8) return run_general_impl<DVec<8>>(*plan, g);
    return TSKB_ERR_UNSUPPORTED;
}

int run_sample_count_stat(const Plan *plan, const StatSpec &spec) {
    std::lock_guard<std::mutex> lock(plan->mu);
    TSKB_CK(cudaSetDevice(plan->device));
    if (spec.K <= 1) return run_impl<IVec<1>>(*plan, spec);
    if (spec.K <= 2) return run_impl<IVec<2>>(*plan, spec);
    if (spec.K <= 4) return run_impl<IVec<4>>(*plan, spec);
    if (spec.K <= 8) return run_impl<IVec<8>>(*plan, spec);
    return TSKB_ERR_UNSUPPORTED;
}

int run_weighted_stat(const Plan *plan, const StatSpec &spec) {
    std::lock_guard<std::mutex> lock(plan->mu);
    TSKB_CK(cudaSetDevice(plan->device));
    if (spec.K <= 1) return run_impl<DVec<1>>(*plan, spec);
    if (spec.K <= 2) return run_impl<DVec<2>>(*plan, spec);
    if (spec.K <= 4) return run_impl<DVec<4>>(*plan, spec);
    if (spec.K <= 8) return run_impl<DVec<8>>(*plan, spec);
    return TSKB_ERR_UNSUPPORTED;
}

int run_trees_at(const Plan *plan, uint64_t nq, const double *positions, const int32_t *tracked,
    uint64_t num_tracked, int32_t *out_parent, int32_t *out_count) {
    std::lock_guard<std::mutex> lock(plan->mu);
    const Plan &P = *plan;
    TSKB_CK(cudaSetDevice(P.device));
    cudaStream_t s = P.stream;
    const uint32_t N = (uint32_t) P.N;
    DevArray<double> d_pos;
    DevArray<int32_t> d_par, d_cnt, d_tr;
    d_pos.upload(positions, nq, s);
    d_par.alloc(nq * N);
    d_cnt.alloc(nq * N);
    const int32_t *tr = P.d_samples.p;
    uint32_t nt = P.num_samples;
    if (tracked != nullptr) {
        d_tr.upload(tracked, num_tracked, s);
        tr = d_tr.p;
        nt = (uint32_t) num_tracked;
    }
    TSKB_CK(cudaMemsetAsync(d_cnt.p, 0, nq * N * sizeof(int32_t), s));
    if (nq * N) {
        k_parent_at<<<grid_for(nq * N, TB), TB, 0, s>>>(d_pos.p, (uint32_t) nq, N, P.coff.p,
            P.csr_left.p, P.csr_right.p, P.csr_parent.p, d_par.p);
        TSKB_CK_LAUNCH();
    }
    if (nq * nt) {
        k_count_at<<<grid_for(nq * nt, TB), TB, 0, s>>>(tr, nt, (uint32_t) nq, N, d_par.p, d_cnt.p);
        TSKB_CK_LAUNCH();
    }
    TSKB_CK(cudaMemcpyAsync(out_parent, d_par.p, nq * N * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    TSKB_CK(cudaMemcpyAsync(out_count, d_cnt.p, nq * N * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    TSKB_CK(cudaStreamSynchronize(s));
    return 0;
}

}  // namespace tskb
