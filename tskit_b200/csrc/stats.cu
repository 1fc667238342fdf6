// stats.cu -- per-call kernels of the sample-count statistics
// (tsk_treeseq_sample_count_stat -> tsk_treeseq_general_stat, c/tskit/trees.c:2035-2220).
//
// Phases of one call (all on the plan's stream):
//   0 weights    sample sets -> 0/1 int32 weight rows per node          (trees.c:2195-2213)
//   1 propagate  ONE launch: addend gather + segmented prefix sum over the node-major addend
//                lists = state[u] over every piece; tiles of a level wait on a completion
//                counter for the levels below                           (trees.c:1317-1327)
//   2 summary    branch: stream the pieces, G = branch_length * f(state), bin G - G_prev into
//                the window holding the piece's left end                (trees.c:1339-1350, 1484-1504)
//                site:   per site, allele states and sum of f           (trees.c:1525-1652)
//   3 finalize   branch: prefix over windows + span-normalise; site: window sums
//                                                                       (trees.c:1753-1762, 1920-1934)
//   5 d2h        result -> host
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "plan.cuh"

namespace tskb {
namespace {

constexpr int TB = 256;

template <int KP>
struct alignas(KP >= 4 ? 16 : 4 * KP) IVec {
    int32_t v[KP];
    __host__ __device__ __forceinline__ IVec operator+(const IVec &o) const {
        IVec r;
#pragma unroll
        for (int k = 0; k < KP; k++) r.v[k] = v[k] + o.v[k];
        return r;
    }
    __host__ __device__ __forceinline__ IVec operator-(const IVec &o) const {
        IVec r;
#pragma unroll
        for (int k = 0; k < KP; k++) r.v[k] = v[k] - o.v[k];
        return r;
    }
};

template <int KP>
__device__ __forceinline__ IVec<KP> ivec_zero() {
    IVec<KP> r;
#pragma unroll
    for (int k = 0; k < KP; k++) r.v[k] = 0;
    return r;
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
    double n[8];            // sample set sizes
    const ColP *cols;       // device [M]
    const double *table;    // device [rows * M] (STAT_TABULATED)
    uint32_t table_rows;
};

// x[i] without dynamic register indexing
template <int KP>
__device__ __forceinline__ double pick(const IVec<KP> &s, int i) {
    int32_t r = s.v[0];
#pragma unroll
    for (int k = 1; k < KP; k++) r = (i == k) ? s.v[k] : r;
    return (double) r;
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
template <int STAT, int KP>
__device__ __forceinline__ double f_eval(const SumP &P, const ColP &c, int m, const IVec<KP> &s) {
    if constexpr (STAT == STAT_DIVERSITY) {
        double n = c.ni, x = pick<KP>(s, c.i);
        return x * (n - x) * c.inv;
    } else if constexpr (STAT == STAT_SEGSITES) {
        double n = c.ni, x = pick<KP>(s, c.i);
        return (x > 0) * (1 - x / n);
    } else if constexpr (STAT == STAT_Y1) {
        double ni = c.ni, xi = pick<KP>(s, c.i);
        double numer = xi * (ni - xi) * (ni - xi - 1);
        return numer * c.inv;
    } else if constexpr (STAT == STAT_DIVERGENCE) {
        return pick<KP>(s, c.i) * (c.nj - pick<KP>(s, c.j)) * c.inv;
    } else if constexpr (STAT == STAT_Y2) {
        double nj = c.nj;
        double xi = pick<KP>(s, c.i), xj = pick<KP>(s, c.j);
        return xi * (nj - xj) * (nj - xj - 1) * c.inv;
    } else if constexpr (STAT == STAT_F2) {
        double ni = c.ni, nj = c.nj;
        double xi = pick<KP>(s, c.i), xj = pick<KP>(s, c.j);
        double numer = xi * (xi - 1) * (nj - xj) * (nj - xj - 1) - xi * (ni - xi) * (nj - xj) * xj;
        return numer * c.inv;
    } else if constexpr (STAT == STAT_RELATEDNESS) {
        double sumx = 0;
#pragma unroll
        for (int k = 0; k < KP; k++) {
            if (k < P.K) sumx += (double) s.v[k] / P.n[k];
        }
        double meanx = sumx / (double) P.K;
        return (pick<KP>(s, c.i) / c.ni - meanx) * (pick<KP>(s, c.j) / c.nj - meanx);
    } else if constexpr (STAT == STAT_RELATEDNESS_NC) {
        return pick<KP>(s, c.i) * pick<KP>(s, c.j) * c.inv;
    } else if constexpr (STAT == STAT_Y3) {
        double numer = pick<KP>(s, c.i) * (c.nj - pick<KP>(s, c.j)) * (c.nk - pick<KP>(s, c.k));
        return numer * c.inv;
    } else if constexpr (STAT == STAT_F3) {
        double ni = c.ni, nj = c.nj, nk = c.nk;
        double xi = pick<KP>(s, c.i), xj = pick<KP>(s, c.j), xk = pick<KP>(s, c.k);
        double numer = xi * (xi - 1) * (nj - xj) * (nk - xk) - xi * (ni - xi) * (nj - xj) * xk;
        return numer * c.inv;
    } else if constexpr (STAT == STAT_F4) {
        double nj = c.nj, nk = c.nk, nl = c.nl;
        double xi = pick<KP>(s, c.i), xj = pick<KP>(s, c.j), xk = pick<KP>(s, c.k),
               xl = pick<KP>(s, c.l);
        double numer = xi * xk * (nj - xj) * (nl - xl) - xi * xl * (nj - xj) * (nk - xk);
        return numer * c.inv;
    } else {  // STAT_TABULATED
        uint32_t cnt = (uint32_t) s.v[0];
        if (cnt >= P.table_rows) cnt = P.table_rows - 1;
        return __ldg(P.table + (size_t) cnt * P.M + m);
    }
}

// branch mode: f(x) + f(total - x) unless polarised (trees.c:1944-1972)
template <int STAT, int KP>
__device__ __forceinline__ double F_branch(const SumP &P, const ColP &c, int m, const IVec<KP> &s,
    const IVec<KP> &totals) {
    double r = f_eval<STAT, KP>(P, c, m, s);
    if (!P.polarised) r += f_eval<STAT, KP>(P, c, m, totals - s);
    return r;
}

// ---------------------------------------------------------------- phase 0

template <int KP>
__global__ void k_set_weights(const int32_t *sets, const uint32_t *set_off, uint32_t K,
    uint32_t total, IVec<KP> *w) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= total) return;
    uint32_t k = upper_bound_dev(set_off, K + 1, j) - 1;
    // a sample may be in several sets (trees.c:2201-2214): distinct columns, no race
    w[sets[j]].v[k] = 1;
}

// ---------------------------------------------------------------- phase 1
// state[u] over every piece (what update_state, trees.c:1317-1327, maintains incrementally):
// a piece's state is the sum of the states of the pieces it references -- its children in the
// tree right of its breakpoint, plus its own INIT piece (the sample weight, trees.c:1406-1415).
// Pieces are processed by height; one cooperative launch of co-resident persistent CTAs covers
// every height.  CTA b takes tiles b, b + G, ...; a tile loads its references, then waits on
// the completion counter until every tile of the lower heights has published its states
// (red.release / ld.acquire at gpu scope), gathers, sums and stores.  Every CTA processes its
// tiles in increasing order and a tile only waits on lower-numbered tiles, so the lowest
// unfinished tile can always run: no deadlock.  Integer sums: exact in any order.

template <int KP>
__device__ __forceinline__ IVec<KP> state_load(const IVec<KP> *p) {
    // states are written by other CTAs of the same launch: read through L2, never L1
    IVec<KP> r;
    if constexpr (KP >= 4) {
        const int4 *q = reinterpret_cast<const int4 *>(p);
#pragma unroll
        for (int c = 0; c < KP / 4; c++) {
            int4 t = __ldcg(q + c);
            r.v[4 * c] = t.x; r.v[4 * c + 1] = t.y; r.v[4 * c + 2] = t.z; r.v[4 * c + 3] = t.w;
        }
    } else if constexpr (KP == 2) {
        int2 t = __ldcg(reinterpret_cast<const int2 *>(p));
        r.v[0] = t.x; r.v[1] = t.y;
    } else {
        r.v[0] = __ldcg(reinterpret_cast<const int *>(p));
    }
    return r;
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

constexpr int PROP_TB = 256;
constexpr int PROP_IPT = PROP_TILE / PROP_TB;
constexpr int PROP_PRE = 3;  // references fetched before the wait; more are rare (multifurcations)
constexpr uint32_t SPIN_LIMIT = 1u << 22;  // ~seconds; a legitimate wait is < the kernel's own run time

template <int KP>
__global__ void __launch_bounds__(PROP_TB) k_propagate(uint32_t ntiles,
    const uint32_t *__restrict__ tile_dep, const uint32_t *__restrict__ pp_piece,
    const uint32_t *__restrict__ pp_off, const uint32_t *__restrict__ refs, IVec<KP> *pval,
    uint32_t *counters, int *error_flag, unsigned long long *trace) {
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        TRACE(0);
        const uint32_t dep = __ldg(tile_dep + tile);
        uint32_t piece[PROP_IPT], o0[PROP_IPT], o1[PROP_IPT], rf[PROP_IPT][PROP_PRE];
        // everything that does not depend on other tiles is fetched before the wait
#pragma unroll
        for (int q = 0; q < PROP_IPT; q++) {
            const uint32_t j = tile * PROP_TILE + q * PROP_TB + threadIdx.x;
            piece[q] = __ldg(pp_piece + j);
            o0[q] = __ldg(pp_off + j);
            o1[q] = __ldg(pp_off + j + 1);
        }
#pragma unroll
        for (int q = 0; q < PROP_IPT; q++) {
#pragma unroll
            for (int i = 0; i < PROP_PRE; i++) {
                rf[q][i] = o0[q] + i < o1[q] ? __ldg(refs + o0[q] + i) : NO_PIECE;
            }
        }
        if (dep > 0) {
            if (threadIdx.x == 0) {
                uint32_t spins = 0;
                while (ld_acquire(counters + 1) < dep) {
                    if (++spins > SPIN_LIMIT || ((spins & 1023u) == 0 && *(volatile int *) error_flag)) {
                        *error_flag = 1;
                        break;
                    }
                }
            }
            __syncthreads();
        }
        TRACE(1);
        IVec<KP> g[PROP_IPT][PROP_PRE];
#pragma unroll
        for (int q = 0; q < PROP_IPT; q++) {
#pragma unroll
            for (int i = 0; i < PROP_PRE; i++) {
                g[q][i] = ivec_zero<KP>();
                if (rf[q][i] != NO_PIECE) g[q][i] = state_load<KP>(pval + rf[q][i]);
            }
        }
        TRACE(2);
#pragma unroll
        for (int q = 0; q < PROP_IPT; q++) {
            IVec<KP> sum = g[q][0];
#pragma unroll
            for (int i = 1; i < PROP_PRE; i++) sum = sum + g[q][i];
            for (uint32_t o = o0[q] + PROP_PRE; o < o1[q]; o++) {
                sum = sum + state_load<KP>(pval + __ldg(refs + o));
            }
            if (piece[q] != NO_PIECE) pval[piece[q]] = sum;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            red_release_add(counters + 1, 1u);
            if (trace != nullptr) trace[(size_t) tile * 4 + 3] = gtime();
        }
    }
}

// INIT pieces hold the node's own sample weight (trees.c:1406-1415)
template <int KP>
__global__ void k_init_pieces(const IVec<KP> *__restrict__ w, const int32_t *__restrict__ rank_node,
    const uint32_t *__restrict__ poff, uint32_t N, IVec<KP> *pval) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < N) pval[poff[r]] = w[rank_node[r]];
}

// ---------------------------------------------------------------- phase 2, branch mode
// Node u contributes  sum over its pieces of  G = branch_length * F(state)  times the overlap
// of the piece with each window.  A piece ends where the next one starts, so this is the sum
// over pieces of (G - G_prev) * |[x, range_right) ^ window|: every piece adds
//   c = G - G_prev   to A[w(x)]  (all windows right of w(x) gain c * their span) and
//   c * (right(w(x)) - x)  to B[w(x)].
// This is the reference's running sum (trees.c:1339-1350, 1484-1504) with the updates of one
// node at one breakpoint telescoped.

constexpr int SUM_IPT = 4;
constexpr int SUM_TILE = TB * SUM_IPT;   // pc_x / pc_bl / pval are padded to whole tiles
constexpr uint32_t LUT_CELLS = 2048;     // uniform cells over the genome -> window index

// window holding x: start from the lookup cell, then walk (windows are sorted; exact for any
// window layout, one step for evenly spaced windows)
__device__ __forceinline__ uint32_t window_of(const double *win, const uint16_t *lut, uint32_t W,
    double x, double inv_cell) {
    uint32_t g = (uint32_t) (x * inv_cell);
    if (g >= LUT_CELLS) g = LUT_CELLS - 1;
    uint32_t u = lut[g];
    while (u > 0 && x < win[u]) u--;
    while (u + 1 < W && x >= win[u + 1]) u++;
    return u;
}

template <int STAT, int KP, bool SMEM>
__global__ void __launch_bounds__(TB, SMEM ? 4 : 2) k_branch_summary(uint32_t ntiles,
    const double *__restrict__ pc_x, const double *__restrict__ pc_bl,
    const IVec<KP> *__restrict__ pval, SumP sp, IVec<KP> totals,
    const double *__restrict__ windows, uint32_t W, double range_right, uint32_t mc, double *gA,
    double *gB) {
    extern __shared__ double smem[];
    double *s_win = smem;                              // [W + 1]
    double *s_bins = smem + (W + 1);                   // [mc][W][2]: A, B interleaved
    uint16_t *s_lut = reinterpret_cast<uint16_t *>(s_bins + (SMEM ? 2 * mc * W : 0));
    const uint32_t m0 = blockIdx.y * mc;
    const uint32_t m1 = min(m0 + mc, (uint32_t) sp.M);
    double inv_cell = 0.0;
    if (SMEM) {
        for (uint32_t i = threadIdx.x; i <= W; i += TB) s_win[i] = windows[i];
        for (uint32_t i = threadIdx.x; i < 2 * mc * W; i += TB) s_bins[i] = 0.0;
        __syncthreads();
        const double cell = (s_win[W] - s_win[0]) / (double) LUT_CELLS;
        inv_cell = 1.0 / cell;
        for (uint32_t g = threadIdx.x; g < LUT_CELLS; g += TB) {
            uint32_t u = upper_bound_dev(s_win, W + 1, s_win[0] + g * cell);
            u = u > 0 ? u - 1 : 0;
            s_lut[g] = (uint16_t) (u < W ? u : W - 1);
        }
        __syncthreads();
    }
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t first = tile * SUM_TILE + threadIdx.x * SUM_IPT;
        double x[SUM_IPT], bl[SUM_IPT + 1];
        IVec<KP> s[SUM_IPT + 1];
        {
            const double2 *qx = reinterpret_cast<const double2 *>(pc_x + first);
            const double2 *qb = reinterpret_cast<const double2 *>(pc_bl + first);
#pragma unroll
            for (int q = 0; q < SUM_IPT / 2; q++) {
                double2 t = qx[q];
                x[2 * q] = t.x; x[2 * q + 1] = t.y;
                t = qb[q];
                bl[2 * q + 1] = t.x; bl[2 * q + 2] = t.y;
            }
            const int4 *qs = reinterpret_cast<const int4 *>(pval + first);
            int flat[SUM_IPT * KP];
#pragma unroll
            for (int q = 0; q < KP; q++) {
                int4 t = qs[q];
                flat[4 * q] = t.x; flat[4 * q + 1] = t.y; flat[4 * q + 2] = t.z; flat[4 * q + 3] = t.w;
            }
#pragma unroll
            for (int q = 0; q < SUM_IPT; q++) {
#pragma unroll
                for (int k = 0; k < KP; k++) s[q + 1].v[k] = flat[q * KP + k];
            }
        }
        // predecessor of the first item (same node unless that item is an INIT piece)
        bool prev_real = false;
        s[0] = ivec_zero<KP>();
        bl[0] = 0.0;
        if (first > 0 && x[0] >= 0.0) {
            prev_real = pc_x[first - 1] >= 0.0;
            bl[0] = pc_bl[first - 1];
            s[0] = pval[first - 1];
        }
        uint32_t wi[SUM_IPT];
        double rem[SUM_IPT];
#pragma unroll
        for (int q = 0; q < SUM_IPT; q++) {
            wi[q] = 0;
            rem[q] = 0.0;
            if (x[q] >= 0.0) {
                uint32_t u;
                double wr;
                if (SMEM) {
                    u = window_of(s_win, s_lut, W, x[q], inv_cell);
                    wr = s_win[u + 1];
                } else {
                    u = upper_bound_dev(windows, W + 1, x[q]);
                    u = u > 0 ? u - 1 : 0;
                    if (u >= W) u = W - 1;
                    wr = windows[u + 1];
                }
                wi[q] = u;
                rem[q] = (wr < range_right ? wr : range_right) - x[q];
            }
        }
        for (uint32_t m = m0; m < m1; m++) {
            const ColP col = sp.cols[m];
            double prevG = prev_real ? bl[0] * F_branch<STAT, KP>(sp, col, m, s[0], totals) : 0.0;
            double *binm = SMEM ? s_bins + 2 * (m - m0) * W : nullptr;
#pragma unroll
            for (int q = 0; q < SUM_IPT; q++) {
                if (x[q] < 0.0) {  // INIT piece (or padding): the node is not in any tree yet
                    prevG = 0.0;
                    continue;
                }
                double G = bl[q + 1] * F_branch<STAT, KP>(sp, col, m, s[q + 1], totals);
                double c = G - prevG;
                prevG = G;
                if (c != 0.0) {
                    if (SMEM) {
                        atomicAdd(binm + 2 * wi[q], c);
                        atomicAdd(binm + 2 * wi[q] + 1, c * rem[q]);
                    } else {
                        atomicAdd(&gA[(size_t) m * W + wi[q]], c);
                        atomicAdd(&gB[(size_t) m * W + wi[q]], c * rem[q]);
                    }
                }
            }
        }
    }
    if (SMEM) {
        __syncthreads();
        const uint32_t cols = m1 - m0;
        for (uint32_t i = threadIdx.x; i < cols * W; i += TB) {
            uint32_t m = m0 + i / W, wdx = i % W;
            double a = s_bins[2 * i], b = s_bins[2 * i + 1];
            if (a != 0.0) atomicAdd(&gA[(size_t) m * W + wdx], a);
            if (b != 0.0) atomicAdd(&gB[(size_t) m * W + wdx], b);
        }
    }
}

// result[w][m] = (sum of A[m][w'] over w' < w) * |window ^ range| + B[m][w], span-normalised
// (trees.c:1920-1934).  One block per column, windows in chunks with a running carry.
__global__ void __launch_bounds__(TB) k_branch_finalize(const double *gA, const double *gB,
    const double *windows, uint32_t W, uint32_t M, double range_left, double range_right,
    int span_normalise, double *result) {
    typedef cub::BlockScan<double, TB> BS;
    __shared__ typename BS::TempStorage tmp;
    __shared__ double s_carry;
    const uint32_t m = blockIdx.x;
    if (threadIdx.x == 0) s_carry = 0.0;
    __syncthreads();
    for (uint32_t base = 0; base < W; base += TB) {
        uint32_t wdx = base + threadIdx.x;
        double a = wdx < W ? gA[(size_t) m * W + wdx] : 0.0;
        double ex, total;
        BS(tmp).ExclusiveSum(a, ex, total);
        double carry = s_carry;
        if (wdx < W) {
            double wl = windows[wdx], wr = windows[wdx + 1];
            double l = wl > range_left ? wl : range_left;
            double r = wr < range_right ? wr : range_right;
            double v = 0.0;
            if (r > l) v = (carry + ex) * (r - l) + gB[(size_t) m * W + wdx];
            if (span_normalise) v /= wr - wl;
            result[(size_t) wdx * M + m] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
}

// ---------------------------------------------------------------- phase 2, site mode

constexpr int MC = 4;  // result columns evaluated per pass over a site's alleles

template <int STAT, int KP>
__global__ void k_site_summary(uint32_t site_lo, uint32_t nsites, const uint32_t *site_moff,
    const uint32_t *site_aoff, const int32_t *mut_src, const uint16_t *mut_allele,
    const uint16_t *mut_alt, const IVec<KP> *pval, IVec<KP> totals, IVec<KP> *scratch, SumP P,
    double *R) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nsites) return;
    uint32_t site = site_lo + t;
    uint32_t a0 = site_aoff[site], na = site_aoff[site + 1] - a0;
    const uint32_t mb = site_moff[site], me = site_moff[site + 1];
    if (na == 2 && me - mb == 1) {
        // one mutation, two alleles: no scratch traffic
        IVec<KP> x = pval[mut_src[mb]];
        IVec<KP> anc = totals - x;
        for (int m = 0; m < P.M; m++) {
            const ColP col = P.cols[m];
            double acc = 0.0;
            if (!P.polarised) acc += f_eval<STAT, KP>(P, col, m, anc);
            acc += f_eval<STAT, KP>(P, col, m, x);
            R[(size_t) m * nsites + t] = acc;
        }
        return;
    }
    scratch[a0] = totals;  // allele 0 starts at total_weight (trees.c:1548)
    for (uint32_t al = 1; al < na; al++) scratch[a0 + al] = ivec_zero<KP>();
    for (uint32_t m = mb; m < me; m++) {
        IVec<KP> x = pval[mut_src[m]];
        scratch[a0 + mut_allele[m]] = scratch[a0 + mut_allele[m]] + x;
        scratch[a0 + mut_alt[m]] = scratch[a0 + mut_alt[m]] - x;
    }
    for (int m = 0; m < P.M; m++) {
        const ColP col = P.cols[m];
        double acc = 0.0;
        for (uint32_t al = P.polarised ? 1 : 0; al < na; al++) {
            acc += f_eval<STAT, KP>(P, col, m, scratch[a0 + al]);
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

constexpr size_t SMEM_BIN_BUDGET = 96 * 1024;

struct CallCtx {
    const Plan *P;
    const StatSpec *sp;
    cudaStream_t s;
    SumP sumP;
    double *d_windows;
    double *d_result;
    uint64_t launches;
};

template <int STAT, int KP>
void launch_branch(CallCtx &c, const IVec<KP> *pval, IVec<KP> totals) {
    const Plan &P = *c.P;
    const uint32_t W = c.sp->W, M = c.sp->M;
    Arena &A = P.arena;
    double *gA = A.get<double>((size_t) 2 * M * W);
    double *gB = gA + (size_t) M * W;
    TSKB_CK(cudaMemsetAsync(gA, 0, (size_t) 2 * M * W * sizeof(double), c.s));
    const uint32_t ntiles = (P.P + SUM_TILE - 1) / SUM_TILE;  // plan arrays are padded to this
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, P.device);
    const size_t win_bytes = (size_t) (W + 1) * sizeof(double) + LUT_CELLS * sizeof(uint16_t);
    const size_t col_bytes = (size_t) 2 * W * sizeof(double);
    if (P.P > 0) {
        if (win_bytes + col_bytes <= SMEM_BIN_BUDGET && W < 65536) {
            uint32_t mc = (uint32_t) std::min<size_t>(M, (SMEM_BIN_BUDGET - win_bytes) / col_bytes);
            uint32_t chunks = (M + mc - 1) / mc;
            size_t smem = win_bytes + mc * col_bytes;
            auto kern = k_branch_summary<STAT, KP, true>;
            TSKB_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
            int per_sm = 1;
            TSKB_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TB, smem));
            uint32_t gx = std::min<uint32_t>(ntiles, (uint32_t) (sms * std::max(per_sm, 1)));
            kern<<<dim3(gx, chunks), TB, smem, c.s>>>(ntiles, P.pc_x.p, P.pc_bl.p, pval, c.sumP, totals,
                c.d_windows, W, P.range_right, mc, gA, gB);
        } else {
            auto kern = k_branch_summary<STAT, KP, false>;
            uint32_t gx = std::min<uint32_t>(ntiles, (uint32_t) sms * 8);
            kern<<<dim3(gx, 1), TB, 0, c.s>>>(ntiles, P.pc_x.p, P.pc_bl.p, pval, c.sumP, totals,
                c.d_windows, W, P.range_right, M, gA, gB);
        }
        TSKB_CK_LAUNCH();
        c.launches++;
    }
    TSKB_CK(cudaEventRecord(P.ev[3], c.s));
    k_branch_finalize<<<M, TB, 0, c.s>>>(gA, gB, c.d_windows, W, M, P.range_left, P.range_right,
        (c.sp->options & TSKB_STAT_SPAN_NORMALISE) ? 1 : 0, c.d_result);
    TSKB_CK_LAUNCH();
    c.launches++;
    TSKB_CK(cudaEventRecord(P.ev[4], c.s));
}

template <int STAT, int KP>
void launch_site(CallCtx &c, const IVec<KP> *pval, IVec<KP> totals) {
    const Plan &P = *c.P;
    const uint32_t W = c.sp->W, M = c.sp->M;
    Arena &A = P.arena;
    const uint32_t nsites = P.site_hi - P.site_lo;
    const uint32_t nsplit = std::max<uint32_t>(1, (592 + W - 1) / W);
    double *partial = A.get<double>((size_t) W * nsplit * M);
    double *R = A.get<double>((size_t) M * std::max<uint32_t>(nsites, 1));
    IVec<KP> *scratch = A.get<IVec<KP>>(P.total_alleles + 1);
    if (nsites) {
        k_site_summary<STAT, KP><<<grid_for(nsites, 128), 128, 0, c.s>>>(P.site_lo, nsites,
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

template <int STAT, int KP>
void launch_summary(CallCtx &c, const IVec<KP> *pval, IVec<KP> totals) {
    if (c.sp->options & TSKB_STAT_BRANCH) {
        launch_branch<STAT, KP>(c, pval, totals);
    } else {
        launch_site<STAT, KP>(c, pval, totals);
    }
}

template <int KP>
int run_impl(const Plan &P, const StatSpec &sp) {
    cudaStream_t s = P.stream;
    const uint32_t K = sp.K, M = sp.M, W = sp.W, N = (uint32_t) P.N;
    Arena &A = P.arena;
    A.reset();
    CallCtx c = {};
    c.P = &P;
    c.sp = &sp;
    c.s = s;
    TSKB_CK(cudaEventRecord(P.ev[0], s));

    // ---- phase 0: weights
    uint64_t total = 0;
    std::vector<uint32_t> h_off(K + 1, 0);
    for (uint32_t k = 0; k < K; k++) {
        total += sp.sizes[k];
        h_off[k + 1] = (uint32_t) total;
    }
    IVec<KP> *w = A.get<IVec<KP>>(N);
    uint32_t *d_off = A.get<uint32_t>(K + 1);
    const int32_t *d_sets = sp.sets;
    if (!sp.sets_on_device) {
        int32_t *tmp_sets = A.get<int32_t>(total);
        TSKB_CK(cudaMemcpyAsync(tmp_sets, sp.sets, total * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        d_sets = tmp_sets;
    }
    TSKB_CK(cudaMemcpyAsync(d_off, h_off.data(), (K + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    c.d_windows = A.get<double>(W + 1);
    TSKB_CK(cudaMemcpyAsync(c.d_windows, sp.windows, (W + 1) * sizeof(double), cudaMemcpyHostToDevice, s));
    TSKB_CK(cudaMemsetAsync(w, 0, (size_t) N * sizeof(IVec<KP>), s));
    k_set_weights<KP><<<grid_for(total, TB), TB, 0, s>>>(d_sets, d_off, K, (uint32_t) total, w);
    TSKB_CK_LAUNCH();
    c.launches++;

    SumP &sumP = c.sumP;
    sumP.K = (int) K;
    sumP.M = (int) M;
    sumP.polarised = (sp.options & TSKB_STAT_POLARISED) ? 1 : 0;
    IVec<KP> totals;
    for (int k = 0; k < KP; k++) totals.v[k] = 0;
    for (uint32_t k = 0; k < K; k++) {
        sumP.n[k] = (double) sp.sizes[k];
        totals.v[k] = (int32_t) sp.sizes[k];
    }
    std::vector<ColP> cols(M);
    {
        for (uint32_t m = 0; m < M; m++) {
            ColP &q = cols[m];
            int32_t t[4] = { (int32_t) (m < K ? m : 0), 0, 0, 0 };
            for (uint32_t a = 0; a < sp.tuple; a++) t[a] = sp.indexes[(size_t) m * sp.tuple + a];
            if (sp.stat_id == STAT_TABULATED) t[0] = 0;
            q.i = t[0]; q.j = t[1]; q.k = t[2]; q.l = t[3];
            q.ni = (double) sp.sizes[t[0]]; q.nj = (double) sp.sizes[t[1]];
            q.nk = (double) sp.sizes[t[2]]; q.nl = (double) sp.sizes[t[3]];
            q.inv = 1.0 / column_denominator(sp.stat_id, q);
        }
        ColP *d_cols = A.get<ColP>(M);
        TSKB_CK(cudaMemcpyAsync(d_cols, cols.data(), M * sizeof(ColP), cudaMemcpyHostToDevice, s));
        sumP.cols = d_cols;
    }
    if (sp.stat_id == STAT_TABULATED) {
        double *d_tab = A.get<double>(sp.table_rows * M);
        TSKB_CK(cudaMemcpyAsync(d_tab, sp.f_table, sp.table_rows * M * sizeof(double),
            cudaMemcpyHostToDevice, s));
        sumP.table = d_tab;
        sumP.table_rows = (uint32_t) sp.table_rows;
    }
    TSKB_CK(cudaEventRecord(P.ev[1], s));

    // ---- phase 1: propagate
    IVec<KP> *pval = A.get<IVec<KP>>((size_t) P.P + 2048);  // summary tiles read whole tiles
    int *d_err = A.get<int>(1);
    TSKB_CK(cudaMemsetAsync(d_err, 0, sizeof(int), s));
    if (N) {
        k_init_pieces<KP><<<grid_for(N, TB), TB, 0, s>>>(w, P.rank_node.p, P.d_poff.p, N, pval);
        TSKB_CK_LAUNCH();
        c.launches++;
    }
    if (P.ntiles) {
        uint32_t *counters = A.get<uint32_t>(2);
        TSKB_CK(cudaMemsetAsync(counters, 0, 2 * sizeof(uint32_t), s));
        unsigned long long *trace = nullptr;
        if (getenv("TSKB_TRACE") != nullptr) {
            trace = A.get<unsigned long long>((size_t) P.ntiles * 4);
            TSKB_CK(cudaMemsetAsync(trace, 0, (size_t) P.ntiles * 4 * sizeof(unsigned long long), s));
            P.stats_trace = trace;
        }
        int per_sm = 1, sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, P.device);
        TSKB_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_propagate<KP>, PROP_TB, 0));
        const uint32_t grid = std::min<uint32_t>(P.ntiles, (uint32_t) (sms * std::max(per_sm, 1)));
        uint32_t a_ntiles = P.ntiles;
        const uint32_t *a_dep = P.tile_dep.p, *a_piece = P.pp_piece.p, *a_off = P.pp_off.p,
                       *a_refs = P.refs.p;
        void *args[] = { &a_ntiles, &a_dep, &a_piece, &a_off, &a_refs, &pval, &counters, &d_err, &trace };
        TSKB_CK(cudaLaunchCooperativeKernel((const void *) k_propagate<KP>, dim3(grid), dim3(PROP_TB),
            args, 0, s));
        c.launches++;
    }
    TSKB_CK(cudaEventRecord(P.ev[2], s));

    // ---- phases 2-3
    c.d_result = sp.result_on_device ? sp.result : A.get<double>((size_t) W * M);
    switch (sp.stat_id) {
        case STAT_DIVERSITY: launch_summary<STAT_DIVERSITY, KP>(c, pval, totals); break;
        case STAT_SEGSITES: launch_summary<STAT_SEGSITES, KP>(c, pval, totals); break;
        case STAT_Y1: launch_summary<STAT_Y1, KP>(c, pval, totals); break;
        case STAT_DIVERGENCE: launch_summary<STAT_DIVERGENCE, KP>(c, pval, totals); break;
        case STAT_Y2: launch_summary<STAT_Y2, KP>(c, pval, totals); break;
        case STAT_F2: launch_summary<STAT_F2, KP>(c, pval, totals); break;
        case STAT_RELATEDNESS: launch_summary<STAT_RELATEDNESS, KP>(c, pval, totals); break;
        case STAT_RELATEDNESS_NC: launch_summary<STAT_RELATEDNESS_NC, KP>(c, pval, totals); break;
        case STAT_Y3: launch_summary<STAT_Y3, KP>(c, pval, totals); break;
        case STAT_F3: launch_summary<STAT_F3, KP>(c, pval, totals); break;
        case STAT_F4: launch_summary<STAT_F4, KP>(c, pval, totals); break;
        case STAT_TABULATED:
            if constexpr (KP == 1) {
                launch_summary<STAT_TABULATED, KP>(c, pval, totals);
                break;
            }
            return TSKB_ERR_UNSUPPORTED;
        default: return TSKB_ERR_BAD_PARAM_VALUE;
    }
    TSKB_CK(cudaEventRecord(P.ev[5], s));
    int h_err = 0;
    TSKB_CK(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (!sp.result_on_device) {
        TSKB_CK(cudaMemcpyAsync(sp.result, c.d_result, (size_t) W * M * sizeof(double),
            cudaMemcpyDeviceToHost, s));
    }
    TSKB_CK(cudaEventRecord(P.ev[6], s));
    TSKB_CK(cudaStreamSynchronize(s));
    float ms = 0;
    for (int q = 0; q < 6; q++) {
        TSKB_CK(cudaEventElapsedTime(&ms, P.ev[q], P.ev[q + 1]));
        P.stats.last_kernel_ms[q] = ms;
    }
    TSKB_CK(cudaEventElapsedTime(&ms, P.ev[0], P.ev[6]));
    P.stats.last_call_ms = ms;
    P.stats.last_launches = c.launches;
    if (h_err) {
        last_error_string() = "propagation wait timed out";
        return TSKB_ERR_CUDA;
    }
    return 0;
}

}  // namespace

int run_sample_count_stat(const Plan *plan, const StatSpec &spec) {
    std::lock_guard<std::mutex> lock(plan->mu);
    TSKB_CK(cudaSetDevice(plan->device));
    if (spec.K <= 1) return run_impl<1>(*plan, spec);
    if (spec.K <= 2) return run_impl<2>(*plan, spec);
    if (spec.K <= 4) return run_impl<4>(*plan, spec);
    if (spec.K <= 8) return run_impl<8>(*plan, spec);
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
