// stats.cu -- per-call kernels of the sample-count statistics
// (tsk_treeseq_sample_count_stat -> tsk_treeseq_general_stat, c/tskit/trees.c:2035-2220).
//
// Phases of one call (all on the plan's stream):
//   0 weights    sample sets -> 0/1 int32 weight rows per node         (trees.c:2195-2213)
//   1 propagate  per dependency level: addend gather + segmented prefix sum over the
//                node-major visit lists = state[u] after every visit   (trees.c:1317-1327)
//   2 summary    branch: per edge diff, change of the running sum      (trees.c:1339-1350, 1425-1474)
//                site:   per site, allele states and sum of f          (trees.c:1525-1652)
//   3 scan       branch: running sum after every diff (prefix sum over diffs)
//   4 windows    integrate / bin into windows, span-normalise          (trees.c:1484-1504, 1753-1762, 1920-1934)
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
constexpr int MC = 4;  // result columns evaluated per pass over an event's visits

template <int KP>
struct alignas(KP >= 4 ? 16 : 4 * KP) IVec {
    int32_t v[KP];
    __host__ __device__ IVec operator+(const IVec &o) const {
        IVec r;
#pragma unroll
        for (int k = 0; k < KP; k++) r.v[k] = v[k] + o.v[k];
        return r;
    }
};

struct SumP {
    int stat;
    int K;
    int M;
    int polarised;
    int skip_zero_bl;  // summaries are finite: a zero branch length contributes exactly 0
    double n[8];
    const int32_t *idx;    // device [M * tuple]
    const double *table;   // device [rows * M] (STAT_TABULATED)
    uint32_t table_rows;
};

// The summary functions, in the reference's exact operation order
// (c/tskit/trees.c:3934-3948, 4221-4264, 4690-4773, 4899-4959, 5177-5291).
// x: the K state values as doubles; P.n: sample set sizes.
__device__ __forceinline__ double f_eval(const SumP &P, const double *x, int m) {
    switch (P.stat) {
        case STAT_DIVERSITY: {
            double n = P.n[m], xm = x[m];
            return xm * (n - xm) / (n * (n - 1));
        }
        case STAT_SEGSITES: {
            double n = P.n[m], xm = x[m];
            return (xm > 0) * (1 - xm / n);
        }
        case STAT_Y1: {
            double ni = P.n[m], xi = x[m];
            double denom = ni * (ni - 1) * (ni - 2);
            double numer = xi * (ni - xi) * (ni - xi - 1);
            return numer / denom;
        }
        case STAT_DIVERGENCE: {
            int i = P.idx[2 * m], j = P.idx[2 * m + 1];
            double ni = P.n[i], nj = P.n[j];
            double denom = ni * (nj - (i == j));
            return x[i] * (nj - x[j]) / denom;
        }
        case STAT_Y2: {
            int i = P.idx[2 * m], j = P.idx[2 * m + 1];
            double ni = P.n[i], nj = P.n[j];
            double xi = x[i], xj = x[j];
            double denom = ni * nj * (nj - 1);
            return xi * (nj - xj) * (nj - xj - 1) / denom;
        }
        case STAT_F2: {
            int i = P.idx[2 * m], j = P.idx[2 * m + 1];
            double ni = P.n[i], nj = P.n[j];
            double xi = x[i], xj = x[j];
            double denom = ni * (ni - 1) * nj * (nj - 1);
            double numer = xi * (xi - 1) * (nj - xj) * (nj - xj - 1)
                           - xi * (ni - xi) * (nj - xj) * xj;
            return numer / denom;
        }
        case STAT_RELATEDNESS: {
            double sumx = 0;
            for (int k = 0; k < P.K; k++) sumx += x[k] / P.n[k];
            double meanx = sumx / (double) P.K;
            int i = P.idx[2 * m], j = P.idx[2 * m + 1];
            double ni = P.n[i], nj = P.n[j];
            return (x[i] / ni - meanx) * (x[j] / nj - meanx);
        }
        case STAT_RELATEDNESS_NC: {
            int i = P.idx[2 * m], j = P.idx[2 * m + 1];
            double ni = P.n[i], nj = P.n[j];
            return x[i] * x[j] / (ni * nj);
        }
        case STAT_Y3: {
            int i = P.idx[3 * m], j = P.idx[3 * m + 1], k = P.idx[3 * m + 2];
            double ni = P.n[i], nj = P.n[j], nk = P.n[k];
            double denom = ni * nj * nk;
            double numer = x[i] * (nj - x[j]) * (nk - x[k]);
            return numer / denom;
        }
        case STAT_F3: {
            int i = P.idx[3 * m], j = P.idx[3 * m + 1], k = P.idx[3 * m + 2];
            double ni = P.n[i], nj = P.n[j], nk = P.n[k];
            double xi = x[i], xj = x[j], xk = x[k];
            double denom = ni * (ni - 1) * nj * nk;
            double numer = xi * (xi - 1) * (nj - xj) * (nk - xk) - xi * (ni - xi) * (nj - xj) * xk;
            return numer / denom;
        }
        case STAT_F4: {
            int i = P.idx[4 * m], j = P.idx[4 * m + 1], k = P.idx[4 * m + 2], l = P.idx[4 * m + 3];
            double ni = P.n[i], nj = P.n[j], nk = P.n[k], nl = P.n[l];
            double xi = x[i], xj = x[j], xk = x[k], xl = x[l];
            double denom = ni * nj * nk * nl;
            double numer = xi * xk * (nj - xj) * (nl - xl) - xi * xl * (nj - xj) * (nk - xk);
            return numer / denom;
        }
        case STAT_TABULATED: {
            uint32_t c = (uint32_t) x[0];
            if (c >= P.table_rows) c = P.table_rows - 1;
            return P.table[(size_t) c * P.M + m];
        }
    }
    return 0.0;
}

// branch mode: f(x) + f(total - x) unless polarised (trees.c:1944-1972)
template <int KP>
__device__ __forceinline__ double F_branch(const SumP &P, const IVec<KP> &c, int m) {
    double x[KP];
#pragma unroll
    for (int k = 0; k < KP; k++) x[k] = (double) c.v[k];
    double r = f_eval(P, x, m);
    if (!P.polarised) {
#pragma unroll
        for (int k = 0; k < KP; k++) x[k] = (k < P.K ? P.n[k] : 0.0) - x[k];
        r += f_eval(P, x, m);
    }
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
// state[u] after every visit = running sum over u's node-major list, whose first
// (INIT) entry is u's own sample weight (trees.c:1406-1415) and whose other
// addends are +-state[child of the diff] (update_state, trees.c:1317-1327).

// library check path: addends materialised, then cub::DeviceScan::InclusiveSumByKey
template <int KP>
__global__ void k_gather_delta(uint32_t begin, uint32_t end, const int32_t *nm_src,
    const uint8_t *nm_flag, const IVec<KP> *val, const IVec<KP> *w, IVec<KP> *delta) {
    uint32_t k = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= end) return;
    int32_t s = nm_src[k];
    uint8_t f = nm_flag[k];
    IVec<KP> x = (f & 2) ? w[s] : val[s];
    if (f & 1) {
#pragma unroll
        for (int q = 0; q < KP; q++) x.v[q] = -x.v[q];
    }
    delta[k - begin] = x;
}

// level 0: nodes that are never a parent; their lists are the INIT entry alone
template <int KP>
__global__ void k_level0(uint32_t end, const int32_t *nm_src, const IVec<KP> *w, IVec<KP> *val) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < end) val[k] = w[nm_src[k]];
}

// Fused addend gather + segmented inclusive scan of one level, single pass with
// decoupled look-back between tiles.  A segment = one node's list; list heads are
// the INIT entries.  (value, head) pairs combine as
//   a (+) b = b.head ? b : (a.value + b.value, a.head).
template <int KP>
struct SegVal {
    IVec<KP> v;
    int head;
};
template <int KP>
struct SegOp {
    __device__ __forceinline__ SegVal<KP> operator()(const SegVal<KP> &a, const SegVal<KP> &b) const {
        if (b.head) return b;
        SegVal<KP> r;
        r.v = a.v + b.v;
        r.head = a.head;
        return r;
    }
};

// descriptors are written by one CTA and read by others in the same launch: bypass L1
template <int KP>
__device__ __forceinline__ SegVal<KP> load_cg(const SegVal<KP> *p) {
    SegVal<KP> r;
    const int *q = reinterpret_cast<const int *>(p);
#pragma unroll
    for (int c = 0; c < KP; c++) r.v.v[c] = __ldcg(q + c);
    r.head = __ldcg(reinterpret_cast<const int *>(&p->head));
    return r;
}
template <int KP>
__device__ __forceinline__ void store_cg(SegVal<KP> *p, const SegVal<KP> &x) {
    int *q = reinterpret_cast<int *>(p);
#pragma unroll
    for (int c = 0; c < KP; c++) __stcg(q + c, x.v.v[c]);
    __stcg(reinterpret_cast<int *>(&p->head), x.head);
}

// look-back descriptor of one tile: status 0 = empty, 1 = aggregate ready, 2 = prefix ready
template <int KP>
struct TileDesc {
    SegVal<KP> agg;
    SegVal<KP> prefix;
    volatile int status;
    int pad[3];
};

constexpr int PROP_TB = 256;
constexpr int PROP_IPT = PROP_TILE / PROP_TB;

template <int KP>
__global__ void __launch_bounds__(PROP_TB) k_propagate_level(uint32_t begin, uint32_t end,
    const int32_t *__restrict__ nm_src, const uint8_t *__restrict__ nm_flag,
    const IVec<KP> *__restrict__ w, IVec<KP> *val, TileDesc<KP> *desc, uint32_t *ticket,
    int *error_flag) {
    typedef cub::BlockScan<SegVal<KP>, PROP_TB> BS;
    __shared__ typename BS::TempStorage tmp;
    __shared__ uint32_t s_tile;
    __shared__ SegVal<KP> s_carry;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = begin + tile * PROP_TILE + threadIdx.x * PROP_IPT;
    SegOp<KP> op;

    // gather addends; thread-local segmented scan
    SegVal<KP> item[PROP_IPT];
#pragma unroll
    for (int q = 0; q < PROP_IPT; q++) {
        uint32_t k = base + q;
        SegVal<KP> x;
#pragma unroll
        for (int c = 0; c < KP; c++) x.v.v[c] = 0;
        x.head = 0;
        if (k < end) {
            int32_t s = nm_src[k];
            uint8_t f = nm_flag[k];
            x.v = (f & 2) ? w[s] : val[s];
            if (f & 1) {
#pragma unroll
                for (int c = 0; c < KP; c++) x.v.v[c] = -x.v.v[c];
            }
            x.head = (f & 2) ? 1 : 0;
        }
        item[q] = q == 0 ? x : op(item[q - 1], x);
    }
    SegVal<KP> thread_agg = item[PROP_IPT - 1];
    SegVal<KP> identity;
#pragma unroll
    for (int c = 0; c < KP; c++) identity.v.v[c] = 0;
    identity.head = 0;
    SegVal<KP> thread_excl, tile_agg;
    BS(tmp).ExclusiveScan(thread_agg, thread_excl, identity, op, tile_agg);

    // decoupled look-back: carry entering this tile
    if (threadIdx.x == 0) {
        SegVal<KP> carry = identity;
        TileDesc<KP> *me = desc + tile;
        if (tile == 0) {
            store_cg(&me->prefix, tile_agg);
            __threadfence();
            me->status = 2;
        } else {
            store_cg(&me->agg, tile_agg);
            __threadfence();
            me->status = 1;
            {
                int32_t p = (int32_t) tile - 1;
                uint32_t spins = 0;
                while (true) {
                    TileDesc<KP> *d = desc + p;
                    int st = d->status;
                    if (st == 0) {
                        if (++spins > (1u << 28)) {
                            *error_flag = 1;
                            break;
                        }
                        continue;
                    }
                    __threadfence();
                    if (st == 2) {
                        carry = op(load_cg(&d->prefix), carry);
                        break;
                    }
                    carry = op(load_cg(&d->agg), carry);
                    if (carry.head || p == 0) break;
                    p--;
                }
            }
            store_cg(&me->prefix, op(carry, tile_agg));
            __threadfence();
            me->status = 2;
        }
        s_carry = carry;
    }
    __syncthreads();
    // values entering this thread: carry (+) thread_excl
    SegVal<KP> in = op(s_carry, thread_excl);
#pragma unroll
    for (int q = 0; q < PROP_IPT; q++) {
        uint32_t k = base + q;
        if (k < end) {
            SegVal<KP> r = op(in, item[q]);
            val[k] = r.v;
        }
    }
}

// ---------------------------------------------------------------- phase 2

constexpr uint32_t CHILD_BIT = 0x80000000u;

// branch mode.  One warp per breakpoint: sum over the diffs at that position of the change of
// the running sum  sum_u branch_length[u] * summary[u]  (update_running_sum,
// trees.c:1339-1350; loops :1425-1474):
//   CHILD entry: +-(time[p] - time[c]) * F(state[c])          (trees.c:1428-1432, 1455-1457)
//   visit entry: bl[u] * (F(state[u] after) - F(state[u] before))  (trees.c:1434-1447, 1460-1473)
template <int KP>
__global__ void k_bp_summary(uint32_t T, const uint32_t *__restrict__ bp_end,
    const uint32_t *__restrict__ em_idx, const double *__restrict__ em_bl,
    const IVec<KP> *__restrict__ val, SumP P, double *B) {
    uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    if (t >= T) return;
    uint32_t j0 = t > 0 ? bp_end[t - 1] : 0, j1 = bp_end[t];
    for (int m0 = 0; m0 < P.M; m0 += MC) {
        double acc[MC];
#pragma unroll
        for (int q = 0; q < MC; q++) acc[q] = 0.0;
        for (uint32_t j = j0 + lane; j < j1; j += 32) {
            uint32_t idx = em_idx[j];
            double bl = em_bl[j];
            uint32_t k = idx & ~CHILD_BIT;
            if (idx & CHILD_BIT) {
                IVec<KP> xc = val[k];
#pragma unroll
                for (int q = 0; q < MC; q++) {
                    if (m0 + q < P.M) acc[q] += bl * F_branch<KP>(P, xc, m0 + q);
                }
            } else if (bl != 0.0 || !P.skip_zero_bl) {
                IVec<KP> xn = val[k], xo = val[k - 1];
#pragma unroll
                for (int q = 0; q < MC; q++) {
                    if (m0 + q < P.M) {
                        acc[q] += bl * (F_branch<KP>(P, xn, m0 + q) - F_branch<KP>(P, xo, m0 + q));
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < MC; q++) {
            double v = acc[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (lane == 0 && m0 + q < P.M) B[(size_t) (m0 + q) * T + t] = v;
        }
    }
}

template <int KP>
__global__ void k_site_summary(uint32_t site_lo, uint32_t nsites, const uint32_t *site_moff,
    const uint32_t *site_aoff, const int32_t *mut_src, const uint16_t *mut_allele,
    const uint16_t *mut_alt, const IVec<KP> *val, IVec<KP> totals, IVec<KP> *scratch, SumP P,
    double *R) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nsites) return;
    uint32_t site = site_lo + t;
    uint32_t a0 = site_aoff[site], na = site_aoff[site + 1] - a0;
    IVec<KP> zero;
#pragma unroll
    for (int k = 0; k < KP; k++) zero.v[k] = 0;
    scratch[a0] = totals;  // allele 0 starts at total_weight (trees.c:1548)
    for (uint32_t al = 1; al < na; al++) scratch[a0 + al] = zero;
    for (uint32_t m = site_moff[site]; m < site_moff[site + 1]; m++) {
        IVec<KP> x = val[mut_src[m]];
        IVec<KP> d = scratch[a0 + mut_allele[m]];
        scratch[a0 + mut_allele[m]] = d + x;
        IVec<KP> e = scratch[a0 + mut_alt[m]];
#pragma unroll
        for (int k = 0; k < KP; k++) e.v[k] -= x.v[k];
        scratch[a0 + mut_alt[m]] = e;
    }
    for (int m0 = 0; m0 < P.M; m0 += MC) {
        double acc[MC];
#pragma unroll
        for (int q = 0; q < MC; q++) acc[q] = 0.0;
        for (uint32_t al = P.polarised ? 1 : 0; al < na; al++) {
            IVec<KP> c = scratch[a0 + al];
            double x[KP];
#pragma unroll
            for (int k = 0; k < KP; k++) x[k] = (double) c.v[k];
#pragma unroll
            for (int q = 0; q < MC; q++) {
                if (m0 + q < P.M) acc[q] += f_eval(P, x, m0 + q);
            }
        }
#pragma unroll
        for (int q = 0; q < MC; q++) {
            if (m0 + q < P.M) R[(size_t) (m0 + q) * nsites + t] = acc[q];
        }
    }
}

// ---------------------------------------------------------------- phase 3
// Deterministic three-step prefix sum over each row of D[M][n]:
// tile sums -> scan of tile sums -> tile-local scan with carry-in.

constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = TB * SCAN_ITEMS;

__global__ void k_tile_reduce(const double *D, uint32_t n, uint32_t ntiles, double *agg) {
    typedef cub::BlockReduce<double, TB> BR;
    __shared__ typename BR::TempStorage tmp;
    uint32_t tile = blockIdx.x, m = blockIdx.y;
    const double *row = D + (size_t) m * n;
    double v[SCAN_ITEMS];
    uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; q++) v[q] = base + q < n ? row[base + q] : 0.0;
    double s = 0;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; q++) s += v[q];
    double tot = BR(tmp).Sum(s);
    if (threadIdx.x == 0) agg[(size_t) m * ntiles + tile] = tot;
}

// one block per row: exclusive scan of the tile sums (in place)
constexpr int AGG_TB = 512;
__global__ void __launch_bounds__(AGG_TB) k_agg_scan(double *agg, uint32_t ntiles) {
    typedef cub::BlockScan<double, AGG_TB> BS;
    __shared__ typename BS::TempStorage tmp;
    __shared__ double carry;
    double *row = agg + (size_t) blockIdx.x * ntiles;
    if (threadIdx.x == 0) carry = 0.0;
    __syncthreads();
    for (uint32_t base = 0; base < ntiles; base += AGG_TB) {
        uint32_t i = base + threadIdx.x;
        double v = i < ntiles ? row[i] : 0.0;
        double ex, total;
        BS(tmp).ExclusiveSum(v, ex, total);
        double c = carry;
        if (i < ntiles) row[i] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
}

__global__ void k_tile_scan(double *D, uint32_t n, uint32_t ntiles, const double *agg) {
    typedef cub::BlockScan<double, TB> BS;
    __shared__ typename BS::TempStorage tmp;
    uint32_t tile = blockIdx.x, m = blockIdx.y;
    double *row = D + (size_t) m * n;
    double v[SCAN_ITEMS];
    uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; q++) v[q] = base + q < n ? row[base + q] : 0.0;
    BS(tmp).InclusiveSum(v, v);
    double c = agg[(size_t) m * ntiles + tile];
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; q++) {
        if (base + q < n) row[base + q] = c + v[q];
    }
}

// ---------------------------------------------------------------- phase 4

// branch: result[w] = sum_t S_t * |[pos_t, pos_{t+1}) ^ window w| over breakpoints t, S_t the
// running sum once every diff at pos_t is applied (trees.c:1484-1504)
__global__ void k_window_branch(const double *windows, uint32_t nsplit, const double *bp_pos,
    uint32_t T, double range_right, const double *S, uint32_t M, double *partial) {
    typedef cub::BlockReduce<double, TB> BR;
    __shared__ typename BR::TempStorage tmp;
    uint32_t w = blockIdx.x, c = blockIdx.y;
    double wl = windows[w], wr = windows[w + 1];
    uint32_t lo = upper_bound_dev(bp_pos, T, wl);
    lo = lo > 0 ? lo - 1 : 0;
    uint32_t hi = lower_bound_dev(bp_pos, T, wr);
    if (hi < lo) hi = lo;
    uint32_t len = hi - lo, per = (len + nsplit - 1) / nsplit;
    uint32_t s0 = lo + c * per, s1 = s0 + per;
    if (s0 > hi) s0 = hi;
    if (s1 > hi) s1 = hi;
    for (uint32_t m = 0; m < M; m++) {
        const double *row = S + (size_t) m * T;
        double sum = 0.0;
        for (uint32_t i = s0 + threadIdx.x; i < s1; i += TB) {
            double p = bp_pos[i];
            double nx = i + 1 < T ? bp_pos[i + 1] : range_right;
            double l = p > wl ? p : wl;
            double r = nx < wr ? nx : wr;
            if (r > l) sum += row[i] * (r - l);
        }
        double tot = BR(tmp).Sum(sum);
        if (threadIdx.x == 0) partial[((size_t) w * nsplit + c) * M + m] = tot;
        __syncthreads();
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

struct Launches {
    uint64_t n = 0;
};

static bool use_cub_propagate() {
    const char *e = getenv("TSKB_PROPAGATE");
    return e != nullptr && strcmp(e, "cub") == 0;
}

template <int KP>
int run_impl(const Plan &P, const StatSpec &sp) {
    cudaStream_t s = P.stream;
    Launches L;
    const uint32_t K = sp.K, M = sp.M, W = sp.W, N = (uint32_t) P.N;
    const bool branch = (sp.options & TSKB_STAT_BRANCH) != 0;
    Arena &A = P.arena;
    A.reset();
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
    double *d_windows = A.get<double>(W + 1);
    TSKB_CK(cudaMemcpyAsync(d_windows, sp.windows, (W + 1) * sizeof(double), cudaMemcpyHostToDevice, s));
    TSKB_CK(cudaMemsetAsync(w, 0, (size_t) N * sizeof(IVec<KP>), s));
    k_set_weights<KP><<<grid_for(total, TB), TB, 0, s>>>(d_sets, d_off, K, (uint32_t) total, w);
    TSKB_CK_LAUNCH();
    L.n++;

    SumP sumP = {};
    sumP.stat = sp.stat_id;
    sumP.K = (int) K;
    sumP.M = (int) M;
    sumP.polarised = (sp.options & TSKB_STAT_POLARISED) ? 1 : 0;
    uint64_t min_size = ~uint64_t(0);
    IVec<KP> totals;
    for (int k = 0; k < KP; k++) totals.v[k] = 0;
    for (uint32_t k = 0; k < K; k++) {
        sumP.n[k] = (double) sp.sizes[k];
        totals.v[k] = (int32_t) sp.sizes[k];
        min_size = std::min<uint64_t>(min_size, sp.sizes[k]);
    }
    sumP.skip_zero_bl = min_size > 3 ? 1 : 0;
    if (sp.tuple > 0) {
        int32_t *d_idx = A.get<int32_t>((size_t) M * sp.tuple);
        TSKB_CK(cudaMemcpyAsync(d_idx, sp.indexes, (size_t) M * sp.tuple * sizeof(int32_t),
            cudaMemcpyHostToDevice, s));
        sumP.idx = d_idx;
    }
    if (sp.stat_id == STAT_TABULATED) {
        double *d_tab = A.get<double>(sp.table_rows * M);
        TSKB_CK(cudaMemcpyAsync(d_tab, sp.f_table, sp.table_rows * M * sizeof(double),
            cudaMemcpyHostToDevice, s));
        sumP.table = d_tab;
        sumP.table_rows = (uint32_t) sp.table_rows;
        bool finite = true;
        for (uint64_t q = 0; q < sp.table_rows * M; q++) finite = finite && std::isfinite(sp.f_table[q]);
        sumP.skip_zero_bl = finite ? 1 : 0;
    }
    TSKB_CK(cudaEventRecord(P.ev[1], s));

    // ---- phase 1: propagate
    IVec<KP> *val = A.get<IVec<KP>>(P.Vn);
    const std::vector<uint32_t> &lb = P.level_begin;
    if (lb[1] > 0) {
        k_level0<KP><<<grid_for(lb[1], TB), TB, 0, s>>>(lb[1], P.nm_src.p, w, val);
        TSKB_CK_LAUNCH();
        L.n++;
    }
    int *d_err = A.get<int>(1);
    TSKB_CK(cudaMemsetAsync(d_err, 0, sizeof(int), s));
    if (use_cub_propagate()) {
        uint32_t max_range = 0;
        for (uint32_t l = 1; l < P.nlevels; l++) max_range = std::max(max_range, lb[l + 1] - lb[l]);
        IVec<KP> *delta = A.get<IVec<KP>>(max_range);
        size_t scan_bytes = 0;
        if (max_range) {
            TSKB_CK(cub::DeviceScan::InclusiveSumByKey(nullptr, scan_bytes, P.nm_key.p, delta, val,
                max_range, ::cuda::std::equal_to<>(), s));
        }
        char *scan_tmp = A.get<char>(scan_bytes);
        for (uint32_t l = 1; l < P.nlevels; l++) {
            uint32_t b0 = lb[l], b1 = lb[l + 1];
            if (b1 == b0) continue;
            k_gather_delta<KP><<<grid_for(b1 - b0, TB), TB, 0, s>>>(b0, b1, P.nm_src.p,
                P.nm_flag.p, val, w, delta);
            L.n++;
            size_t bytes = scan_bytes;
            TSKB_CK(cub::DeviceScan::InclusiveSumByKey(scan_tmp, bytes, P.nm_key.p + b0, delta,
                val + b0, b1 - b0, ::cuda::std::equal_to<>(), s));
        }
    } else {
        const uint32_t ntiles = P.level_tile0[P.nlevels];
        TileDesc<KP> *desc = A.get<TileDesc<KP>>(ntiles + 1);
        uint32_t *ticket = A.get<uint32_t>(P.nlevels + 1);
        TSKB_CK(cudaMemsetAsync(desc, 0, (size_t) (ntiles + 1) * sizeof(TileDesc<KP>), s));
        TSKB_CK(cudaMemsetAsync(ticket, 0, (P.nlevels + 1) * sizeof(uint32_t), s));
        for (uint32_t l = 1; l < P.nlevels; l++) {
            uint32_t b0 = lb[l], b1 = lb[l + 1];
            if (b1 == b0) continue;
            uint32_t tiles = P.level_tile0[l + 1] - P.level_tile0[l];
            k_propagate_level<KP><<<tiles, PROP_TB, 0, s>>>(b0, b1, P.nm_src.p, P.nm_flag.p, w,
                val, desc + P.level_tile0[l], ticket + l, d_err);
            L.n++;
        }
    }
    TSKB_CK_LAUNCH();
    TSKB_CK(cudaEventRecord(P.ev[2], s));

    // ---- phases 2-4
    const uint32_t nsplit = std::max<uint32_t>(1, (592 + W - 1) / W);
    double *partial = A.get<double>((size_t) W * nsplit * M);
    double *d_result = sp.result_on_device ? sp.result : A.get<double>((size_t) W * M);
    const int span_norm = (sp.options & TSKB_STAT_SPAN_NORMALISE) ? 1 : 0;
    if (branch) {
        const uint32_t T = P.T;
        double *B = A.get<double>((size_t) M * std::max<uint32_t>(T, 1));
        if (T) {
            k_bp_summary<KP><<<grid_for((size_t) T * 32, 128), 128, 0, s>>>(T, P.bp_end.p,
                P.em_idx.p, P.em_bl.p, val, sumP, B);
            TSKB_CK_LAUNCH();
            L.n++;
        }
        TSKB_CK(cudaEventRecord(P.ev[3], s));
        if (T) {
            uint32_t ntiles = (T + SCAN_TILE - 1) / SCAN_TILE;
            double *agg = A.get<double>((size_t) M * ntiles);
            k_tile_reduce<<<dim3(ntiles, M), TB, 0, s>>>(B, T, ntiles, agg);
            k_agg_scan<<<M, AGG_TB, 0, s>>>(agg, ntiles);
            k_tile_scan<<<dim3(ntiles, M), TB, 0, s>>>(B, T, ntiles, agg);
            TSKB_CK_LAUNCH();
            L.n += 3;
        }
        TSKB_CK(cudaEventRecord(P.ev[4], s));
        k_window_branch<<<dim3(W, nsplit), TB, 0, s>>>(d_windows, nsplit, P.bp_pos.p, T,
            P.range_right, B, M, partial);
        TSKB_CK_LAUNCH();
        L.n++;
    } else {
        const uint32_t nsites = P.site_hi - P.site_lo;
        double *R = A.get<double>((size_t) M * std::max<uint32_t>(nsites, 1));
        IVec<KP> *scratch = A.get<IVec<KP>>(P.total_alleles + 1);
        if (nsites) {
            k_site_summary<KP><<<grid_for(nsites, 128), 128, 0, s>>>(P.site_lo, nsites,
                P.site_moff.p, P.site_aoff.p, P.mut_src.p, P.mut_allele.p, P.mut_alt.p, val,
                totals, scratch, sumP, R);
            TSKB_CK_LAUNCH();
            L.n++;
        }
        TSKB_CK(cudaEventRecord(P.ev[3], s));
        TSKB_CK(cudaEventRecord(P.ev[4], s));
        k_window_site<<<dim3(W, nsplit), TB, 0, s>>>(d_windows, nsplit, P.site_pos.p, P.site_lo,
            nsites, R, M, partial);
        TSKB_CK_LAUNCH();
        L.n++;
    }
    k_window_final<<<grid_for((size_t) W * M, TB), TB, 0, s>>>(partial, d_windows, W, nsplit, M,
        span_norm, d_result);
    TSKB_CK_LAUNCH();
    L.n++;
    TSKB_CK(cudaEventRecord(P.ev[5], s));
    int h_err = 0;
    TSKB_CK(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (!sp.result_on_device) {
        TSKB_CK(cudaMemcpyAsync(sp.result, d_result, (size_t) W * M * sizeof(double),
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
    P.stats.last_launches = L.n;
    if (h_err) {
        last_error_string() = "propagation look-back timed out";
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
