// plan.cu -- staging: tables -> HBM, and construction of the replay plan (see plan.cuh).
//
// Everything here runs once per tree sequence (the counterpart of
// tsk_treeseq_init, c/tskit/trees.c:455-545).  Sorting / compaction / scans use
// CUB device primitives; the plan-specific kernels are below.
#include <cub/cub.cuh>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>

#include "plan.cuh"

namespace tskb {

std::string &last_error_string() {
    static thread_local std::string s;
    return s;
}

Plan::~Plan() {
    arena.destroy();
    for (auto &e : ev) {
        if (e) cudaEventDestroy(e);
    }
    if (stream) cudaStreamDestroy(stream);
}

uint64_t Plan::device_bytes() const {
    return time.bytes() + d_samples.bytes() + coff.bytes() + csr_left.bytes() + csr_right.bytes()
           + csr_parent.bytes() + ev_pos.bytes() + ev_child.bytes() + ev_sign.bytes()
           + voff.bytes() + q_off.bytes() + refs.bytes() + q_bp0.bytes() + q_bp1.bytes() + bp_pos.bytes()
           + q_bl.bytes() + q_eff.bytes() + q_node.bytes() + node_first_bp.bytes() + tile_dep.bytes() + d_sample_index.bytes() + pm_off.bytes() + pm_left.bytes()
           + pm_right.bytes() + pm_pmax.bytes() + pm_child.bytes() + rank_node.bytes() + level.bytes() + site_pos.bytes() + site_moff.bytes()
           + site_aoff.bytes() + mut_node.bytes() + mut_src.bytes() + mut_allele.bytes()
           + mut_alt.bytes() + so_slot.bytes() + so_bp0.bytes() + so_bp1.bytes() + so_bl.bytes() + q_x0.bytes() + q_x1.bytes();
}

namespace {

constexpr int TB = 256;

// ------------------------------------------------------------------ kernels

__global__ void k_adjacent_diff(const uint32_t *off, uint32_t n, uint32_t *out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = off[i + 1] - off[i];
}

__global__ void k_iota(uint32_t *out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

// order-preserving map of an IEEE double onto uint64 (node times may be negative)
__device__ inline uint64_t ordered_bits(double x) {
    uint64_t b = (uint64_t) __double_as_longlong(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

// seed candidates: prefix of the insertion index with left <= range_left;
// keep those still alive at range_left (right > range_left)
__global__ void k_seed_flags(const int32_t *I, uint32_t n, const double *er, double a,
    uint8_t *flag, uint64_t *key, const int32_t *ep, const double *time) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) {
        int32_t e = I[j];
        flag[j] = er[e] > a;
        key[j] = ordered_bits(time[ep[e]]);
    }
}

__global__ void k_build_ins(const int32_t *seed_edges, uint32_t n_seed, const int32_t *I_rest,
    uint32_t n_ins, const double *el, double a, int32_t *ins_edge, double *ins_pos) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_ins) {
        int32_t e = j < n_seed ? seed_edges[j] : I_rest[j - n_seed];
        ins_edge[j] = e;
        double l = el[e];
        ins_pos[j] = l > a ? l : a;
    }
}

__global__ void k_build_rem(const int32_t *O_slice, uint32_t n_rem, const double *er,
    int32_t *rem_edge, double *rem_pos) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_rem) {
        int32_t e = O_slice[k];
        rem_edge[k] = e;
        rem_pos[k] = er[e];
    }
}

// Event index = position in the reference's processing order: at one
// breakpoint all removals precede all insertions (trees.c:1425-1474).
__global__ void k_merge_events(const int32_t *ins_edge, const double *ins_pos, uint32_t n_ins,
    const int32_t *rem_edge, const double *rem_pos, uint32_t n_rem, const int32_t *ep,
    const int32_t *ec, const double *time, double *ev_pos, int32_t *ev_child, int32_t *ev_parent,
    int8_t *ev_sign, double *ev_sbl) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_ins + n_rem) return;
    uint32_t idx;
    int32_t e;
    double pos;
    int8_t sign;
    if (t < n_ins) {
        pos = ins_pos[t];
        e = ins_edge[t];
        idx = t + upper_bound_dev(rem_pos, n_rem, pos);
        sign = 1;
    } else {
        uint32_t k = t - n_ins;
        pos = rem_pos[k];
        e = rem_edge[k];
        idx = k + lower_bound_dev(ins_pos, n_ins, pos);
        sign = -1;
    }
    int32_t p = ep[e], c = ec[e];
    ev_pos[idx] = pos;
    ev_child[idx] = c;
    ev_parent[idx] = p;
    ev_sign[idx] = sign;
    double bl = time[p] - time[c];  // trees.c:1456
    ev_sbl[idx] = sign > 0 ? bl : -bl;
}

// canonical index order check: removals old parent first, insertions young parent first
__global__ void k_check_order(const double *ev_pos, const int8_t *ev_sign,
    const int32_t *ev_parent, const double *time, uint32_t nev, int *bad) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= nev) return;
    if (ev_pos[i] == ev_pos[i + 1] && ev_sign[i] == ev_sign[i + 1]) {
        double t0 = time[ev_parent[i]], t1 = time[ev_parent[i + 1]];
        if ((ev_sign[i] < 0 && t0 < t1) || (ev_sign[i] > 0 && t0 > t1)) *bad = 1;
    }
}

__global__ void k_csr_keys(const int32_t *I, const int32_t *ec, uint32_t E, uint32_t *key,
    uint32_t *val) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < E) {
        int32_t e = I[j];
        key[j] = (uint32_t) ec[e];
        val[j] = (uint32_t) e;
    }
}

__global__ void k_csr_gather(const uint32_t *edge, uint32_t E, const double *el, const double *er,
    const int32_t *ep, double *csr_left, double *csr_right, int32_t *csr_parent) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < E) {
        uint32_t e = edge[j];
        csr_left[j] = el[e];
        csr_right[j] = er[e];
        csr_parent[j] = ep[e];
    }
}

// out[q] = lower_bound(sorted_keys, q) for q in [0, nq)
__global__ void k_offsets(const uint32_t *sorted_keys, uint32_t n, uint32_t nq, uint32_t *out) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) out[q] = lower_bound_dev(sorted_keys, n, q);
}

// parent of u through an edge spanning ACROSS x: left < x < right (and x inside the range)
__device__ inline int32_t span_parent(int32_t u, double x, double a, const uint32_t *coff,
    const double *csr_left, const double *csr_right, const int32_t *csr_parent) {
    if (!(x > a)) return -1;
    uint32_t lo = coff[u], hi = coff[u + 1];
    uint32_t k = lower_bound_dev(csr_left + lo, hi - lo, x);  // edges with left < x
    if (k == 0) return -1;
    uint32_t e = lo + k - 1;
    return csr_right[e] > x ? csr_parent[e] : -1;
}

constexpr uint32_t MAX_CHAIN = 1u << 22;

__global__ void k_chain_count(const int32_t *ev_parent, const double *ev_pos, uint32_t nev,
    double a, const uint32_t *coff, const double *csr_left, const double *csr_right,
    const int32_t *csr_parent, uint32_t *count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nev) return;
    double x = ev_pos[i];
    int32_t u = ev_parent[i];
    uint32_t c = 0;
    while (u != -1 && c < MAX_CHAIN) {
        c++;
        u = span_parent(u, x, a, coff, csr_left, csr_right, csr_parent);
    }
    count[i] = c;
}

__global__ void k_chain_fill(const int32_t *ev_parent, const double *ev_pos, uint32_t nev,
    double a, const uint32_t *coff, const double *csr_left, const double *csr_right,
    const int32_t *csr_parent, const double *time, const uint32_t *voff, int32_t *vis_node,
    double *vis_bl, uint32_t *vis_ev) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nev) return;
    double x = ev_pos[i];
    int32_t u = ev_parent[i];
    uint32_t j = voff[i], end = voff[i + 1];
    while (u != -1 && j < end) {
        int32_t v = span_parent(u, x, a, coff, csr_left, csr_right, csr_parent);
        vis_node[j] = u;
        vis_ev[j] = i;
        vis_bl[j] = v == -1 ? 0.0 : time[v] - time[u];
        j++;
        u = v;
    }
}

__global__ void k_relax_levels(const int32_t *ep, const int32_t *ec, uint32_t E, uint32_t *level,
    int *changed) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    uint32_t lc = level[ec[e]] + 1;
    int32_t p = ep[e];
    if (level[p] < lc) {
        atomicMax(&level[p], lc);
        *changed = 1;
    }
}

__global__ void k_scatter_rank(const uint32_t *rank_node, uint32_t N, uint32_t *rank) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < N) rank[rank_node[r]] = r;
}

__global__ void k_bp_flags(const double *ev_pos, uint32_t nev, uint32_t *flag) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nev) flag[i] = (i + 1 == nev || ev_pos[i] != ev_pos[i + 1]) ? 1u : 0u;
}

// Entries in event-major order: for event i, its CHILD entry (the edge's own child gets a new
// piece: its branch appears / disappears) at e = voff[i] + i, then its visits bottom-up.
__global__ void k_entry_keys_child(const int32_t *ev_child, const uint32_t *voff,
    const uint32_t *rank, uint32_t nev, uint32_t *key, uint32_t *em_ev) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nev) return;
    uint32_t e = voff[i] + i;
    key[e] = rank[ev_child[i]];
    em_ev[e] = i;
}
__global__ void k_entry_keys_visit(const int32_t *vis_node, const uint32_t *vis_ev,
    const uint32_t *rank, uint32_t V, uint32_t *key, uint32_t *em_ev) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= V) return;
    uint32_t i = vis_ev[j];
    uint32_t e = j + i + 1;
    key[e] = rank[vis_node[j]];
    em_ev[e] = i;
}

// node-major position k of every entry; an entry ends its piece when the next entry of the
// order belongs to another node or another breakpoint
__global__ void k_entry_ends(const uint32_t *sorted_e, const uint32_t *sorted_key,
    const uint32_t *em_ev, const uint32_t *ev_bp, uint32_t Ve, uint32_t *inv, uint32_t *endflag) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Ve) return;
    uint32_t e = sorted_e[k];
    inv[e] = k;
    uint32_t end = 1;
    if (k + 1 < Ve && sorted_key[k + 1] == sorted_key[k]) {
        end = ev_bp[em_ev[sorted_e[k + 1]]] != ev_bp[em_ev[e]];
    }
    endflag[k] = end;
}

__global__ void k_fill_f64(double *out, size_t n, double v) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}
__global__ void k_fill_u32(uint32_t *out, size_t n, uint32_t v) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}

// Pieces: one per (node, breakpoint) group of node-major entries, after the node's INIT piece.
__global__ void k_piece_fill(const uint32_t *sorted_e, const uint32_t *sorted_key,
    const uint32_t *em_ev, const uint32_t *endflag, const uint32_t *endscan, const uint32_t *voff,
    const int8_t *ev_sign, const double *ev_sbl, const double *ev_pos, const double *vis_bl,
    uint32_t Ve, double *pc_x, double *pc_bl, uint32_t *piece_rank) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Ve || !endflag[k]) return;
    uint32_t e = sorted_e[k], r = sorted_key[k], i = em_ev[e];
    bool child = e == voff[i] + i;
    // the piece's branch length is the one in force after the LAST diff of the breakpoint that
    // touches the node: its own insertion if there is one (trees.c:1455-1457), 0 after its
    // removal (trees.c:1432), else the branch across x
    double bl = child ? (ev_sign[i] > 0 ? ev_sbl[i] : 0.0) : vis_bl[e - i - 1];
    uint32_t p = endscan[k] + r + 1;
    pc_x[p] = ev_pos[i];
    pc_bl[p] = bl;
    piece_rank[p] = r;
}

// Does the reference "update" the node at this breakpoint (a visit in some walk, or the removal of
// its own edge) or is the node only the child of an inserted edge?  The branch-mode allele
// frequency spectrum credits a node from its last update (tsk_treeseq_update_branch_afs,
// trees.c:3650-3697: `last_update` is not refreshed for the child of an insertion).
__global__ void k_piece_updates(const uint32_t *sorted_e, const uint32_t *sorted_key,
    const uint32_t *em_ev, const uint32_t *endscan, const uint32_t *voff, const int8_t *ev_sign,
    uint32_t Ve, uint8_t *piece_upd) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Ve) return;
    uint32_t e = sorted_e[k], r = sorted_key[k], i = em_ev[e];
    const bool child = e == voff[i] + i;
    if (!child || ev_sign[i] < 0) piece_upd[endscan[k] + r + 1] = 1;
}

__global__ void k_piece_init(const uint32_t *noff, const uint32_t *endscan, uint32_t Ve,
    uint32_t ends_total, uint32_t N, double *pc_x, double *pc_bl, uint32_t *piece_rank,
    uint32_t *poff) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > N) return;
    uint32_t k = r < N ? noff[r] : Ve;
    uint32_t p = (k < Ve ? endscan[k] : ends_total) + r;
    poff[r] = p;
    if (r < N) {
        pc_x[p] = -1.0;
        pc_bl[p] = 0.0;
        piece_rank[p] = r;
    }
}

// ---- parent-major CSR: edges sorted by (parent, left), running max of right per parent
__global__ void k_pm_keys(const uint32_t *coff, uint32_t N, const double *csr_left, uint32_t E,
    uint64_t *key, uint32_t *val, int32_t *child_of) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= E) return;
    key[j] = ordered_bits(csr_left[j]);
    val[j] = j;
    child_of[j] = (int32_t) (upper_bound_dev(coff, N + 1, j) - 1);
}
__global__ void k_gather_parent(const uint32_t *perm, const int32_t *src, uint32_t n, uint32_t *out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t) src[perm[i]];
}
__global__ void k_pm_gather(const uint32_t *perm, uint32_t E, const double *csr_left,
    const double *csr_right, const int32_t *child_of, double *left, double *right, int32_t *child) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    uint32_t j = perm[i];
    left[i] = csr_left[j];
    right[i] = csr_right[j];
    child[i] = child_of[j];
}
struct MaxOp {
    __device__ __forceinline__ double operator()(double a, double b) const { return a > b ? a : b; }
};

// Children of a piece's node in the tree right of the piece's breakpoint (edges with
// left <= x < right), as the pieces holding their state there; plus the node's own INIT piece
// when it is a sample (its own weight, trees.c:1406-1415).
template <bool FILL>
__global__ void k_piece_children(uint32_t P, const double *pc_x, const uint32_t *piece_rank,
    const int32_t *rank_node, const uint32_t *rank, const uint32_t *node_is_sample,
    const uint32_t *poff, const uint32_t *pm_off, const double *pm_left, const double *pm_right,
    const double *pm_pmax, const int32_t *pm_child, uint32_t *cnt, const uint32_t *ch_off,
    uint32_t *refs) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const double x = pc_x[p];
    uint32_t c = 0;
    if (x >= 0.0) {  // INIT pieces have no references: they are written from the weights
        const uint32_t r = piece_rank[p];
        const int32_t u = rank_node[r];
        uint32_t base = FILL ? ch_off[p] : 0;
        if (node_is_sample[u]) {
            if (FILL) refs[base] = poff[r];
            c++;
        }
        uint32_t lo = pm_off[u], hi = pm_off[u + 1];
        uint32_t k = upper_bound_dev(pm_left + lo, hi - lo, x);  // edges with left <= x
        for (uint32_t j = lo + k; j-- > lo;) {
            if (!(pm_pmax[j] > x)) break;  // nothing further left reaches x
            if (pm_right[j] > x) {
                if (FILL) {
                    uint32_t rc = rank[pm_child[j]];
                    uint32_t q0 = poff[rc], n = poff[rc + 1] - q0;
                    // the child's last piece starting at or before x (its INIT piece starts at -1)
                    refs[base + c] = q0 + upper_bound_dev(pc_x + q0, n, x) - 1;
                }
                c++;
            }
        }
    }
    if (!FILL) cnt[p] = c;
}

// height of a piece = 1 + max height of the pieces it references, INIT pieces not counted (they
// are available before the sweep starts); relaxed to a fixed point
__global__ void k_relax_height(uint32_t P, const uint32_t *ch_off, const uint32_t *refs,
    const double *pc_x, uint32_t *height, int *changed) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    uint32_t h = 0;
    for (uint32_t i = ch_off[p]; i < ch_off[p + 1]; i++) {
        uint32_t q = refs[i];
        if (pc_x[q] >= 0.0) h = max(h, height[q] + 1);
    }
    if (h != height[p]) {
        height[p] = h;
        *changed = 1;
    }
}

__global__ void k_height_keys(uint32_t P, const uint8_t *needed, const uint32_t *height, uint32_t *key,
    uint32_t *val, uint32_t init_key) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    key[p] = needed[p] ? height[p] : init_key;  // INIT and unneeded pieces sort last and are dropped
    val[p] = p;
}

// A piece without a branch above it (a root, or a detached node) is no other piece's child over
// its span and adds nothing to a branch statistic: it is computed only if a mutation sits on it
// (site mode reads state[mutation.node], trees.c:1744-1763).
__global__ void k_needed_branch(uint32_t P, const double *pc_x, const double *pc_bl, int keep_all,
    uint8_t *needed) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) needed[p] = pc_x[p] >= 0.0 && (keep_all || pc_bl[p] != 0.0);
}

// node mode: the node of every piece in processing order, and where each node's pieces begin
__global__ void k_piece_nodes(uint32_t P, const double *pc_x, const uint32_t *piece_rank,
    const int32_t *rank_node, const uint32_t *perm, uint32_t npp, int32_t *q_node) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P || pc_x[p] < 0.0) return;
    const uint32_t j = perm[p];
    if (j < npp) q_node[j] = rank_node[piece_rank[p]];
}
__global__ void k_node_first_bp(uint32_t N, const uint32_t *poff, const int32_t *rank_node,
    const double *pc_x, const double *bp_pos, uint32_t T, uint32_t *node_first_bp) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    const uint32_t p = poff[r] + 1;  // the piece after the INIT piece, if it is this node's
    node_first_bp[rank_node[r]] = p < poff[r + 1] ? lower_bound_dev(bp_pos, T, pc_x[p]) : T;
}
__global__ void k_needed_mutation(uint32_t Mu, const int32_t *mut_src, const double *pc_x, uint8_t *needed) {
    uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < Mu && pc_x[mut_src[m]] >= 0.0) needed[mut_src[m]] = 1;
}

// processing order: pieces by (height, secondary order); every height padded to whole tiles.
// perm[p] = processing position (= state slot) of node-major piece p.
__global__ void k_order_fill(const uint32_t *sorted_piece, const uint32_t *sorted_h, uint32_t nreal,
    const uint32_t *lvl_sorted_begin, const uint32_t *lvl_padded_begin, const uint32_t *cnt,
    const double *pc_x, const double *pc_bl, uint32_t P, const double *bp_pos, uint32_t T,
    const uint8_t *piece_upd, uint32_t *q_bp0, uint32_t *q_bp1, uint32_t *q_eff, double *q_bl,
    uint32_t *q_cnt, uint32_t *perm) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nreal) return;
    uint32_t h = sorted_h[i], p = sorted_piece[i];
    uint32_t j = lvl_padded_begin[h] + (i - lvl_sorted_begin[h]);
    perm[p] = j;
    q_bp0[j] = lower_bound_dev(bp_pos, T, pc_x[p]);
    // the piece ends where the node's next piece starts (the next node's list opens with an
    // INIT marker, x = -1), else at the end of the range (breakpoint index T)
    q_bp1[j] = (p + 1 < P && pc_x[p + 1] >= 0.0) ? lower_bound_dev(bp_pos, T, pc_x[p + 1]) : T;
    q_bl[j] = pc_bl[p];
    q_cnt[j] = cnt[p];
    // where the reference's last_update of the node stands when this piece begins: the piece's own
    // breakpoint if the node is updated there, else the start of the node's previous piece
    // (necessarily an update: the node was parentless), else the start of the range
    uint32_t eff = q_bp0[j];
    if (!piece_upd[p]) eff = pc_x[p - 1] >= 0.0 ? lower_bound_dev(bp_pos, T, pc_x[p - 1]) : 0xffffffffu;
    q_eff[j] = eff;
}

__global__ void k_bp_pos(const double *ev_pos, const uint32_t *ev_bp, uint32_t nev, uint32_t T,
    double range_right, double *bp_pos) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) bp_pos[T] = range_right;
    if (i < nev && (i + 1 == nev || ev_pos[i] != ev_pos[i + 1])) bp_pos[ev_bp[i]] = ev_pos[i];
}

// INIT pieces: the sample's weight slot, or the shared zero slot for a node that is not a sample
__global__ void k_init_perm(uint32_t N, const uint32_t *poff, const int32_t *rank_node,
    const int32_t *sample_index, uint32_t npp, uint32_t n, uint32_t *perm) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    int32_t si = sample_index[rank_node[r]];
    perm[poff[r]] = si >= 0 ? npp + (uint32_t) si : npp + n;
}

__global__ void k_refs_reorder(uint32_t nreal, const uint32_t *sorted_piece, const uint32_t *perm,
    const uint32_t *q_off, const uint32_t *ch_off, const uint32_t *refs, uint32_t *refs2) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nreal) return;
    uint32_t p = sorted_piece[i];
    uint32_t j = perm[p];
    uint32_t o = q_off[j], n = q_off[j + 1] - o, s0 = ch_off[p];
    // ascending state slot: consecutive pieces of a node then list the same children in the same
    // order, so the lanes of a warp gather each slot from neighbouring addresses
    for (uint32_t k = 0; k < n; k++) {
        uint32_t v = perm[refs[s0 + k]];
        uint32_t i = k;
        while (i > 0 && refs2[o + i - 1] > v) {
            refs2[o + i] = refs2[o + i - 1];
            i--;
        }
        refs2[o + i] = v;
    }
}

__global__ void k_x_keys(uint32_t P, const double *pc_x, uint64_t *key, uint32_t *val) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    key[p] = ordered_bits(pc_x[p]);
    val[p] = p;
}

__global__ void k_height_keys_of(uint32_t P, const uint32_t *piece, const uint8_t *needed,
    const uint32_t *height, uint32_t *key, uint32_t init_key) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    uint32_t p = piece[i];
    key[i] = needed[p] ? height[p] : init_key;
}

__global__ void k_level_as_height(uint32_t P, const uint32_t *piece_rank, const int32_t *rank_node,
    const uint32_t *level, uint32_t *height) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) height[p] = level[rank_node[piece_rank[p]]];
}

__global__ void k_translate(int32_t *a, uint32_t n, const uint32_t *perm) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = (int32_t) perm[a[i]];
}

__global__ void k_flag_samples(const int32_t *samples, uint32_t n, uint32_t *flag) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[samples[i]] = 1;
}

// state[mutation.node] at the site's tree = the node's last piece starting at or before the
// site position (trees.c:1744-1763); INIT pieces start at -1
__global__ void k_mut_src(const int32_t *mut_site, const int32_t *mut_node, uint32_t Mu,
    const double *site_pos, const uint32_t *rank, const uint32_t *poff, const double *pc_x,
    int32_t *mut_src) {
    uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= Mu) return;
    double x = site_pos[mut_site[m]];
    uint32_t r = rank[mut_node[m]];
    uint32_t lo = poff[r], n = poff[r + 1] - lo;
    mut_src[m] = (int32_t) (lo + upper_bound_dev(pc_x + lo, n, x) - 1);
}

// ------------------------------------------------------------ allele codes
// Allele strings -> small integer codes, one thread per site (replaces the memcmp loops of
// get_allele_weights, trees.c:1557-1596): allele 0 is the ancestral state, new derived states are
// numbered in order of first appearance among the site's mutations (genotypes.c:533-580).  A
// mutation's "alt" allele -- the one its node's samples are taken from -- is its parent mutation's
// allele, or the ancestral state.  Quadratic in the mutations of one site, which are few.
__device__ __forceinline__ bool bytes_equal(const char *a, uint64_t na, const char *b, uint64_t nb) {
    if (na != nb) return false;
    for (uint64_t i = 0; i < na; i++) {
        if (a[i] != b[i]) return false;
    }
    return true;
}

__global__ void k_allele_codes(uint32_t S, const uint32_t *__restrict__ moff, const char *__restrict__ anc,
    const uint64_t *__restrict__ anc_off, const char *__restrict__ der, const uint64_t *__restrict__ der_off,
    const int32_t *__restrict__ mparent, uint16_t *m_allele, uint16_t *m_alt, uint32_t *nalleles, int *overflow) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= S) return;
    const uint32_t m0 = moff[j], m1 = moff[j + 1];
    const char *a = anc + anc_off[j];
    const uint64_t na = anc_off[j + 1] - anc_off[j];
    uint32_t count = 1;
    for (uint32_t m = m0; m < m1; m++) {
        const char *d = der + der_off[m];
        const uint64_t nd = der_off[m + 1] - der_off[m];
        int k = bytes_equal(d, nd, a, na) ? 0 : -1;
        // an earlier mutation of the site that introduced the same string
        for (uint32_t q = m0; k < 0 && q < m; q++) {
            if (bytes_equal(d, nd, der + der_off[q], der_off[q + 1] - der_off[q])) k = (int) m_allele[q];
        }
        if (k < 0) k = (int) count++;
        if (count > 65000u) {
            *overflow = 1;
            return;
        }
        m_allele[m] = (uint16_t) k;
        const int32_t pm = mparent != nullptr ? mparent[m] : -1;
        m_alt[m] = pm >= (int32_t) m0 && pm < (int32_t) m ? m_allele[pm] : (uint16_t) 0;
    }
    nalleles[j] = count;
}

// ------------------------------------------------------------ table integrity
// The checks of tsk_table_collection_check_integrity (c/tskit/tables.c:10362-10640, 10894-10930)
// that this path relies on -- every id it later uses as an index, every interval, the node time
// ordering -- in the reference's order: nodes, edges, sites, mutations, indexes; within a table the
// first failing row, within a row the first failing check.  key = table << 56 | row << 8 | check.
enum : uint32_t { CK_NODE_TIME = 0, CK_NULL_PARENT, CK_PARENT_BOUNDS, CK_NULL_CHILD, CK_CHILD_BOUNDS,
    CK_COORDS, CK_LEFT, CK_RIGHT, CK_INTERVAL, CK_TIME_ORDER, CK_SITE_POS, CK_SITE_UNSORTED, CK_SITE_DUP,
    CK_MUT_SITE, CK_MUT_NODE, CK_MUT_PARENT_BOUNDS, CK_MUT_PARENT_EQUAL, CK_MUT_PARENT_AFTER,
    CK_MUT_PARENT_SITE, CK_MUT_UNSORTED, CK_INDEX };

__device__ __forceinline__ void ck_fail(unsigned long long *key, uint32_t table, uint64_t row, uint32_t check) {
    atomicMin(key, ((unsigned long long) table << 56) | ((unsigned long long) row << 8) | check);
}

__global__ void k_check_nodes_edges(uint32_t N, const double *time, uint32_t E, const double *el,
    const double *er, const int32_t *ep, const int32_t *ec, const int32_t *I, const int32_t *O, double L,
    unsigned long long *key) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < N && !isfinite(time[j])) ck_fail(key, 0, j, CK_NODE_TIME);
    if (j >= E) return;
    const int32_t p = ep[j], c = ec[j];
    const double l = el[j], r = er[j];
    if (p == -1) ck_fail(key, 1, j, CK_NULL_PARENT);
    else if (p < 0 || (uint32_t) p >= N) ck_fail(key, 1, j, CK_PARENT_BOUNDS);
    else if (c == -1) ck_fail(key, 1, j, CK_NULL_CHILD);
    else if (c < 0 || (uint32_t) c >= N) ck_fail(key, 1, j, CK_CHILD_BOUNDS);
    else if (!(isfinite(l) && isfinite(r))) ck_fail(key, 1, j, CK_COORDS);
    else if (l < 0) ck_fail(key, 1, j, CK_LEFT);
    else if (r > L) ck_fail(key, 1, j, CK_RIGHT);
    else if (l >= r) ck_fail(key, 1, j, CK_INTERVAL);
    else if (time[c] >= time[p]) ck_fail(key, 1, j, CK_TIME_ORDER);
    if (I[j] < 0 || (uint32_t) I[j] >= E || O[j] < 0 || (uint32_t) O[j] >= E) ck_fail(key, 4, j, CK_INDEX);
}

__global__ void k_check_sites_mutations(uint32_t N, uint32_t S, const double *pos, double L, uint32_t Mu,
    const int32_t *msite, const int32_t *mnode, const int32_t *mparent, unsigned long long *key) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < S) {
        const double x = pos[j];
        if (!isfinite(x) || x < 0 || x >= L) ck_fail(key, 2, j, CK_SITE_POS);
        else if (j > 0 && pos[j - 1] == x) ck_fail(key, 2, j, CK_SITE_DUP);
        else if (j > 0 && pos[j - 1] > x) ck_fail(key, 2, j, CK_SITE_UNSORTED);
    }
    if (j >= Mu) return;
    const int32_t st = msite[j], nd = mnode[j], pm = mparent != nullptr ? mparent[j] : -1;
    if (st < 0 || (uint32_t) st >= S) ck_fail(key, 3, j, CK_MUT_SITE);
    else if (nd < 0 || (uint32_t) nd >= N) ck_fail(key, 3, j, CK_MUT_NODE);
    else if (pm < -1 || (pm >= 0 && (uint32_t) pm >= Mu)) ck_fail(key, 3, j, CK_MUT_PARENT_BOUNDS);
    else if (pm == (int32_t) j) ck_fail(key, 3, j, CK_MUT_PARENT_EQUAL);
    else if (pm > (int32_t) j) ck_fail(key, 3, j, CK_MUT_PARENT_AFTER);
    else if (pm >= 0 && msite[pm] != st) ck_fail(key, 3, j, CK_MUT_PARENT_SITE);
    else if (j > 0 && msite[j - 1] > st) ck_fail(key, 3, j, CK_MUT_UNSORTED);
}

int check_code(unsigned long long key) {
    static const int codes[] = { -210 /*TIME_NONFINITE*/, -300 /*NULL_PARENT*/, -202, -301 /*NULL_CHILD*/, -202,
        -211 /*GENOME_COORDS_NONFINITE*/, -310 /*LEFT_LESS_ZERO*/, -309 /*RIGHT_GREATER_SEQ_LENGTH*/,
        -307 /*BAD_EDGE_INTERVAL*/, -306 /*BAD_NODE_TIME_ORDERING*/, -402 /*BAD_SITE_POSITION*/,
        -400 /*UNSORTED_SITES*/, -401 /*DUPLICATE_SITE_POSITION*/, -205 /*SITE_OUT_OF_BOUNDS*/, -202,
        -206 /*MUTATION_OUT_OF_BOUNDS*/, -501 /*MUTATION_PARENT_EQUAL*/, -502 /*MUTATION_PARENT_AFTER_CHILD*/,
        -500 /*MUTATION_PARENT_DIFFERENT_SITE*/, -504 /*UNSORTED_MUTATIONS*/, -203 /*EDGE_OUT_OF_BOUNDS*/ };
    return codes[key & 0xffu];
}

// sum of a uint32 array in 64 bits (the 32-bit scans below must not wrap unnoticed)
struct U32ToU64 {
    __host__ __device__ uint64_t operator()(uint32_t x) const { return (uint64_t) x; }
};

// ------------------------------------------------------------ CUB wrappers

struct Temp {
    DevArray<char> buf;  // pool-backed during a build (common.cuh: alloc_stream)
    void *need(size_t bytes) {
        if (bytes > buf.n) buf.alloc(bytes + (bytes >> 2) + 1024);
        return buf.p;
    }
};

uint64_t sum_u64(Temp &tmp, const uint32_t *p, size_t n, cudaStream_t s) {
    if (n == 0) return 0;
    DevArray<uint64_t> out;
    out.alloc(1);
    cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t *> it(p, U32ToU64());
    size_t bytes = 0;
    TSKB_CK(cub::DeviceReduce::Sum(nullptr, bytes, it, out.p, (int64_t) n, s));
    TSKB_CK(cub::DeviceReduce::Sum(tmp.need(bytes), bytes, it, out.p, (int64_t) n, s));
    uint64_t h = 0;
    TSKB_CK(cudaMemcpyAsync(&h, out.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    TSKB_CK(cudaStreamSynchronize(s));
    return h;
}

// ---- table upload: pageable host columns -> HBM through pinned staging buffers
// cudaMemcpy from pageable memory is staged by the driver on one thread (5-8 GB/s here); the borrowed
// table columns cannot be pinned in place cheaply.  A few host threads copy 4 MB chunks into their own
// pinned buffers (kept for the life of the process) and launch the DMA of each on their own stream, so
// the host copies run in parallel and overlap the transfers.
struct UploadJob {
    void *dst;
    const void *src;
    size_t bytes;
};

struct PinnedPool {
    static constexpr size_t CHUNK = size_t(4) << 20;
    static constexpr int MAX_THREADS = 8;
    std::mutex mu;      // guards buf
    std::mutex use_mu;  // one transfer at a time uses the buffers
    char *buf[MAX_THREADS][2] = {};
    char *get(int t, int b) {
        std::lock_guard<std::mutex> lock(mu);
        if (buf[t][b] == nullptr) {
            if (cudaHostAlloc((void **) &buf[t][b], CHUNK, cudaHostAllocDefault) != cudaSuccess) {
                cudaGetLastError();
                buf[t][b] = nullptr;
            }
        }
        return buf[t][b];
    }
};

PinnedPool &pinned_pool() {
    static PinnedPool *p = new PinnedPool();  // never destroyed: the buffers outlive every plan
    return *p;
}

void staged_upload(int device, const std::vector<UploadJob> &jobs, cudaStream_t fallback) {
    struct Chunk { char *dst; const char *src; size_t n; };
    std::vector<Chunk> chunks;
    size_t total = 0;
    for (const UploadJob &j : jobs) {
        for (size_t o = 0; o < j.bytes; o += PinnedPool::CHUNK) {
            chunks.push_back({ (char *) j.dst + o, (const char *) j.src + o, std::min(PinnedPool::CHUNK, j.bytes - o) });
        }
        total += j.bytes;
    }
    int nt = (int) std::min<size_t>({ (size_t) PinnedPool::MAX_THREADS, std::max<size_t>(1, std::thread::hardware_concurrency() / 2),
        std::max<size_t>(1, chunks.size() / 4) });
    if (getenv("TSKB_UPLOAD_THREADS") != nullptr) nt = std::max(1, std::min(atoi(getenv("TSKB_UPLOAD_THREADS")), (int) PinnedPool::MAX_THREADS));
    if (total < (size_t(16) << 20) || nt <= 1) {  // small tables: the plain copy
        for (const UploadJob &j : jobs) {
            if (j.bytes) TSKB_CK(cudaMemcpyAsync(j.dst, j.src, j.bytes, cudaMemcpyHostToDevice, fallback));
        }
        return;
    }
    std::lock_guard<std::mutex> pool_lock(pinned_pool().use_mu);  // the buffers are shared by every plan
    TSKB_CK(cudaStreamSynchronize(fallback));  // the destinations were allocated in this stream's order
    std::atomic<size_t> next{ 0 };
    std::atomic<int> failed{ 0 };
    auto worker = [&](int t) {
        cudaStream_t st = nullptr;
        cudaEvent_t ev[2] = { nullptr, nullptr };
        char *b[2] = { pinned_pool().get(t, 0), pinned_pool().get(t, 1) };
        bool ok = cudaSetDevice(device) == cudaSuccess && b[0] != nullptr && b[1] != nullptr
                  && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess
                  && cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming) == cudaSuccess
                  && cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming) == cudaSuccess;
        bool used[2] = { false, false };
        int cur = 0;
        while (ok) {
            const size_t i = next.fetch_add(1);
            if (i >= chunks.size()) break;
            if (used[cur]) ok = cudaEventSynchronize(ev[cur]) == cudaSuccess;
            if (!ok) break;
            memcpy(b[cur], chunks[i].src, chunks[i].n);
            ok = cudaMemcpyAsync(chunks[i].dst, b[cur], chunks[i].n, cudaMemcpyHostToDevice, st) == cudaSuccess
                 && cudaEventRecord(ev[cur], st) == cudaSuccess;
            used[cur] = true;
            cur ^= 1;
        }
        if (st != nullptr && cudaStreamSynchronize(st) != cudaSuccess) ok = false;
        if (!ok) failed.store(1);
        for (auto &e : ev) {
            if (e) cudaEventDestroy(e);
        }
        if (st) cudaStreamDestroy(st);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(worker, t);
    worker(0);
    for (auto &x : th) x.join();
    if (failed.load()) {
        // a worker could not stage (no pinned memory left ...): the plain copy of everything
        cudaGetLastError();
        for (const UploadJob &j : jobs) {
            if (j.bytes) TSKB_CK(cudaMemcpyAsync(j.dst, j.src, j.bytes, cudaMemcpyHostToDevice, fallback));
        }
    }
}

}  // namespace

// Large results back to the caller's pageable buffer: the mirror image of staged_upload.  Each
// thread copies 4 MB chunks device -> its pinned buffer -> destination, so the host-side copies
// (and the page faults of a freshly allocated result array) run in parallel.
void staged_download(int device, void *dst, const void *src, size_t bytes, cudaStream_t stream) {
    TSKB_CK(cudaStreamSynchronize(stream));  // the producers of src ran on this stream
    const size_t CH = PinnedPool::CHUNK;
    const size_t nchunks = (bytes + CH - 1) / CH;
    int nt = (int) std::min<size_t>({ (size_t) PinnedPool::MAX_THREADS, std::max<size_t>(1, std::thread::hardware_concurrency() / 2),
        std::max<size_t>(1, nchunks / 4) });
    if (getenv("TSKB_UPLOAD_THREADS") != nullptr) nt = std::max(1, std::min(atoi(getenv("TSKB_UPLOAD_THREADS")), (int) PinnedPool::MAX_THREADS));
    if (bytes < (size_t(16) << 20) || nt <= 1) {
        if (bytes) TSKB_CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
        TSKB_CK(cudaStreamSynchronize(stream));
        return;
    }
    std::lock_guard<std::mutex> pool_lock(pinned_pool().use_mu);
    std::atomic<size_t> next{ 0 };
    std::atomic<int> failed{ 0 };
    auto worker = [&](int t) {
        cudaStream_t st = nullptr;
        char *b = pinned_pool().get(t, 0);
        bool ok = cudaSetDevice(device) == cudaSuccess && b != nullptr
                  && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
        while (ok) {
            const size_t i = next.fetch_add(1);
            if (i >= nchunks) break;
            const size_t off = i * CH, n = std::min(CH, bytes - off);
            ok = cudaMemcpyAsync(b, (const char *) src + off, n, cudaMemcpyDeviceToHost, st) == cudaSuccess
                 && cudaStreamSynchronize(st) == cudaSuccess;
            if (ok) memcpy((char *) dst + off, b, n);
        }
        if (!ok) failed.store(1);
        if (st) cudaStreamDestroy(st);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(worker, t);
    worker(0);
    for (auto &x : th) x.join();
    if (failed.load()) {
        cudaGetLastError();
        TSKB_CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
        TSKB_CK(cudaStreamSynchronize(stream));
    }
}

namespace {

template <typename K, typename Vt>
void sort_pairs(Temp &tmp, const K *kin, K *kout, const Vt *vin, Vt *vout, uint32_t n, int end_bit,
    cudaStream_t s) {
    if (n == 0) return;
    size_t bytes = 0;
    TSKB_CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, 0, end_bit, s));
    void *t = tmp.need(bytes);
    TSKB_CK(cub::DeviceRadixSort::SortPairs(t, bytes, kin, kout, vin, vout, n, 0, end_bit, s));
}

uint32_t host_upper_bound_indexed(const double *col, const int32_t *order, uint64_t n, double x) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = lo + ((hi - lo) >> 1);
        if (col[order[mid]] <= x) lo = mid + 1; else hi = mid;
    }
    return (uint32_t) lo;
}
uint32_t host_lower_bound_indexed(const double *col, const int32_t *order, uint64_t n, double x) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = lo + ((hi - lo) >> 1);
        if (col[order[mid]] < x) lo = mid + 1; else hi = mid;
    }
    return (uint32_t) lo;
}

}  // namespace

// ------------------------------------------------------------------ build

Plan *build_plan(const tskb_tables_t *t, int device, double range_left, double range_right,
    uint32_t options) {
    auto t_start = std::chrono::steady_clock::now();
    TSKB_CK(cudaSetDevice(device));
    std::unique_ptr<Plan> plan(new Plan());
    Plan &P = *plan;
    P.device = device;
    TSKB_CK(cudaStreamCreateWithFlags(&P.stream, cudaStreamNonBlocking));
    for (auto &e : P.ev) TSKB_CK(cudaEventCreate(&e));
    cudaStream_t s = P.stream;
    // temporaries of the build: stream-ordered pool, kept warm until the build is over
    struct PoolScope {
        cudaMemPool_t pool = nullptr;
        cudaStream_t stream;
        explicit PoolScope(int device, cudaStream_t st) : stream(st) {
            if (getenv("TSKB_NO_POOL") == nullptr && cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
                alloc_stream() = st;
            } else {
                cudaGetLastError();
                pool = nullptr;
            }
        }
        ~PoolScope() {
            alloc_stream() = nullptr;
            if (pool != nullptr) {
                cudaStreamSynchronize(stream);
                unsigned long long keep = 0;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
                cudaMemPoolTrimTo(pool, 0);  // what the plan does not hold goes back to the device
            }
        }
    } pool_scope(device, s);
    P.N = t->num_nodes;
    P.E = t->num_edges;
    P.S = t->num_sites;
    P.Mu = t->num_mutations;
    P.L = t->sequence_length;
    P.range_left = range_left;
    P.range_right = range_right;
    P.time_uncalibrated = t->time_uncalibrated;
    const uint32_t N = (uint32_t) P.N, E = (uint32_t) P.E;
    const double a = range_left, b = range_right;
    // TSKB_TIMING=1: wall time of every staging phase, to stderr (the stream is drained at each mark)
    const bool timing = getenv("TSKB_TIMING") != nullptr;
    auto t_mark = t_start;
    auto mark = [&](const char *what) {
        if (!timing) return;
        cudaStreamSynchronize(s);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "tskb init: %-28s %8.2f ms\n", what,
            std::chrono::duration<double, std::milli>(now - t_mark).count());
        t_mark = now;
    };
    mark("context + stream");

    // samples / sample_index_map exactly as init_nodes (trees.c:404-453)
    P.sample_index_map.assign(N, -1);
    for (uint32_t u = 0; u < N; u++) {
        if (t->node_flags[u] & 1u) {
            P.sample_index_map[u] = (int32_t) P.samples.size();
            P.samples.push_back((int32_t) u);
        }
    }
    P.num_samples = (uint32_t) P.samples.size();
    P.d_samples.upload(P.samples.data(), P.samples.size(), s);
    for (uint32_t u = 0; u < N; u++) P.has_negative_time |= t->node_time[u] < 0.0;

    Temp tmp;
    DevArray<double> el, er;
    DevArray<int32_t> ep, ec, dI, dO;
    DevArray<int32_t> d_msite;
    P.time.alloc(N);
    el.alloc(E); er.alloc(E); ep.alloc(E); ec.alloc(E); dI.alloc(E); dO.alloc(E);
    P.site_pos.alloc(P.S); P.mut_node.alloc(P.Mu); d_msite.alloc(P.Mu);
    staged_upload(device, {
        { P.time.p, t->node_time, (size_t) N * sizeof(double) },
        { el.p, t->edge_left, (size_t) E * sizeof(double) }, { er.p, t->edge_right, (size_t) E * sizeof(double) },
        { ep.p, t->edge_parent, (size_t) E * sizeof(int32_t) }, { ec.p, t->edge_child, (size_t) E * sizeof(int32_t) },
        { dI.p, t->edge_insertion_order, (size_t) E * sizeof(int32_t) },
        { dO.p, t->edge_removal_order, (size_t) E * sizeof(int32_t) },
        { P.site_pos.p, t->site_position, (size_t) P.S * sizeof(double) },
        { P.mut_node.p, t->mutation_node, (size_t) P.Mu * sizeof(int32_t) },
        { d_msite.p, t->mutation_site, (size_t) P.Mu * sizeof(int32_t) } }, s);

    mark("host loops + table upload");
    // ---- table integrity, before anything is used as an index on the host or the device
    {
        if (!(P.L > 0)) throw (int) -701;  // TSK_ERR_BAD_SEQUENCE_LENGTH
        if (P.N >= 0x7fffffffull || P.E >= 0x7fffffffull || P.S >= 0x7fffffffull || P.Mu >= 0x7fffffffull) {
            throw (int) TSKB_ERR_UNSUPPORTED;  // tsk_id_t rows
        }
        DevArray<unsigned long long> key;
        DevArray<int32_t> d_mpar;
        key.alloc(1);
        TSKB_CK(cudaMemsetAsync(key.p, 0xff, sizeof(unsigned long long), s));
        if (t->mutation_parent != nullptr) d_mpar.upload(t->mutation_parent, P.Mu, s);
        const uint32_t S = (uint32_t) P.S, Mu = (uint32_t) P.Mu;
        if (std::max(N, E)) {
            k_check_nodes_edges<<<grid_for(std::max(N, E), TB), TB, 0, s>>>(N, P.time.p, E, el.p, er.p, ep.p,
                ec.p, dI.p, dO.p, P.L, key.p);
            TSKB_CK_LAUNCH();
        }
        if (std::max(S, Mu)) {
            k_check_sites_mutations<<<grid_for(std::max(S, Mu), TB), TB, 0, s>>>(N, S, P.site_pos.p, P.L, Mu,
                d_msite.p, P.mut_node.p, d_mpar.p, key.p);
            TSKB_CK_LAUNCH();
        }
        unsigned long long h_key = 0;
        TSKB_CK(cudaMemcpyAsync(&h_key, key.p, sizeof(h_key), cudaMemcpyDeviceToHost, s));
        TSKB_CK(cudaStreamSynchronize(s));
        if (h_key != ~0ull) throw (int) check_code(h_key);
        // ragged columns (host only): offsets start at 0 and never decrease (TSK_ERR_BAD_OFFSET)
        for (int col = 0; col < 2; col++) {
            const uint64_t *off = col == 0 ? t->site_ancestral_state_offset : t->mutation_derived_state_offset;
            const uint64_t rows = col == 0 ? P.S : P.Mu;
            if (rows == 0) continue;
            if (off == nullptr || off[0] != 0) throw (int) -200;
            for (uint64_t j = 0; j < rows; j++) {
                if (off[j + 1] < off[j]) throw (int) -200;
            }
        }
    }

    mark("integrity checks");
    // ---- event lists (host binary searches over the borrowed host columns)
    const uint32_t i0 = host_upper_bound_indexed(t->edge_left, t->edge_insertion_order, E, a);
    const uint32_t i1 = host_lower_bound_indexed(t->edge_left, t->edge_insertion_order, E, b);
    const uint32_t r0 = host_upper_bound_indexed(t->edge_right, t->edge_removal_order, E, a);
    const uint32_t r1 = host_lower_bound_indexed(t->edge_right, t->edge_removal_order, E, b);
    DevArray<int32_t> seed_sorted;
    const int32_t *seed_edges = dI.p;
    uint32_t n_seed = i0;
    if (a > 0 && i0 > 0) {
        // tree at range_left: alive edges, inserted bottom-up (time[parent] ascending)
        DevArray<uint8_t> flag;
        DevArray<uint64_t> key, key_sel, key_out;
        DevArray<int32_t> sel;
        DevArray<uint32_t> nsel;
        flag.alloc(i0); key.alloc(i0); key_sel.alloc(i0); key_out.alloc(i0); sel.alloc(i0);
        seed_sorted.alloc(i0); nsel.alloc(1);
        k_seed_flags<<<grid_for(i0, TB), TB, 0, s>>>(dI.p, i0, er.p, a, flag.p, key.p, ep.p,
            P.time.p);
        TSKB_CK_LAUNCH();
        size_t bytes = 0;
        TSKB_CK(cub::DeviceSelect::Flagged(nullptr, bytes, dI.p, flag.p, sel.p, nsel.p, i0, s));
        TSKB_CK(cub::DeviceSelect::Flagged(tmp.need(bytes), bytes, dI.p, flag.p, sel.p, nsel.p, i0, s));
        TSKB_CK(cub::DeviceSelect::Flagged(nullptr, bytes, key.p, flag.p, key_sel.p, nsel.p, i0, s));
        TSKB_CK(cub::DeviceSelect::Flagged(tmp.need(bytes), bytes, key.p, flag.p, key_sel.p, nsel.p, i0, s));
        TSKB_CK(cudaMemcpyAsync(&n_seed, nsel.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        TSKB_CK(cudaStreamSynchronize(s));
        sort_pairs(tmp, key_sel.p, key_out.p, sel.p, seed_sorted.p, n_seed, 64, s);
        TSKB_CK(cudaStreamSynchronize(s));
        seed_edges = seed_sorted.p;
    }
    const uint32_t n_ins = n_seed + (i1 > i0 ? i1 - i0 : 0);
    const uint32_t n_rem = r1 > r0 ? r1 - r0 : 0;
    const uint32_t nev = n_ins + n_rem;
    P.nev = nev;

    DevArray<int32_t> ev_parent;
    DevArray<double> ev_sbl;
    {
        DevArray<int32_t> ins_edge, rem_edge;
        DevArray<double> ins_pos, rem_pos;
        ins_edge.alloc(n_ins); ins_pos.alloc(n_ins); rem_edge.alloc(n_rem); rem_pos.alloc(n_rem);
        if (n_ins) {
            k_build_ins<<<grid_for(n_ins, TB), TB, 0, s>>>(seed_edges, n_seed, dI.p + i0, n_ins,
                el.p, a, ins_edge.p, ins_pos.p);
            TSKB_CK_LAUNCH();
        }
        if (n_rem) {
            k_build_rem<<<grid_for(n_rem, TB), TB, 0, s>>>(dO.p + r0, n_rem, er.p, rem_edge.p,
                rem_pos.p);
            TSKB_CK_LAUNCH();
        }
        P.ev_pos.alloc(nev); P.ev_child.alloc(nev); P.ev_sign.alloc(nev); ev_sbl.alloc(nev);
        ev_parent.alloc(nev);
        if (nev) {
            k_merge_events<<<grid_for(nev, TB), TB, 0, s>>>(ins_edge.p, ins_pos.p, n_ins,
                rem_edge.p, rem_pos.p, n_rem, ep.p, ec.p, P.time.p, P.ev_pos.p, P.ev_child.p,
                ev_parent.p, P.ev_sign.p, ev_sbl.p);
            TSKB_CK_LAUNCH();
        }
        TSKB_CK(cudaStreamSynchronize(s));
    }
    {
        DevArray<int> bad;
        bad.alloc(1);
        TSKB_CK(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
        if (nev > 1) {
            k_check_order<<<grid_for(nev, TB), TB, 0, s>>>(P.ev_pos.p, P.ev_sign.p, ev_parent.p,
                P.time.p, nev, bad.p);
            TSKB_CK_LAUNCH();
        }
        int h_bad = 0;
        TSKB_CK(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        TSKB_CK(cudaStreamSynchronize(s));
        if (h_bad) {
            throw (int) TSKB_ERR_BAD_INDEX_ORDER;
        }
    }

    mark("event lists + order check");
    // ---- child-major CSR over all edges, sorted by (child, left)
    P.coff.alloc(N + 1);
    P.csr_left.alloc(E); P.csr_right.alloc(E); P.csr_parent.alloc(E);
    {
        DevArray<uint32_t> kin, kout, vin, vout;
        kin.alloc(E); kout.alloc(E); vin.alloc(E); vout.alloc(E);
        if (E) {
            k_csr_keys<<<grid_for(E, TB), TB, 0, s>>>(dI.p, ec.p, E, kin.p, vin.p);
            TSKB_CK_LAUNCH();
            sort_pairs(tmp, kin.p, kout.p, vin.p, vout.p, E, (int) std::max(1u, ceil_log2(N)), s);
            k_csr_gather<<<grid_for(E, TB), TB, 0, s>>>(vout.p, E, el.p, er.p, ep.p, P.csr_left.p,
                P.csr_right.p, P.csr_parent.p);
            TSKB_CK_LAUNCH();
        }
        k_offsets<<<grid_for(N + 1, TB), TB, 0, s>>>(kout.p, E, N + 1, P.coff.p);
        TSKB_CK_LAUNCH();
        TSKB_CK(cudaStreamSynchronize(s));
    }

    mark("child-major CSR");
    // ---- chains: count, scan, fill
    P.voff.alloc(nev + 1);
    uint32_t V = 0;
    {
        DevArray<uint32_t> cnt;
        cnt.alloc(nev + 1);
        TSKB_CK(cudaMemsetAsync(cnt.p, 0, (nev + 1) * sizeof(uint32_t), s));
        if (nev) {
            k_chain_count<<<grid_for(nev, TB), TB, 0, s>>>(ev_parent.p, P.ev_pos.p, nev, a,
                P.coff.p, P.csr_left.p, P.csr_right.p, P.csr_parent.p, cnt.p);
            TSKB_CK_LAUNCH();
        }
        // the scan below is 32-bit: refuse (instead of wrapping) when the visits do not fit
        if (sum_u64(tmp, cnt.p, (size_t) nev + 1, s) + nev + N >= 0xffffffffull) {
            throw (int) TSKB_ERR_UNSUPPORTED;  // 32-bit piece indexes; shard the genome instead
        }
        size_t bytes = 0;
        TSKB_CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt.p, P.voff.p, nev + 1, s));
        TSKB_CK(cub::DeviceScan::ExclusiveSum(tmp.need(bytes), bytes, cnt.p, P.voff.p, nev + 1, s));
        TSKB_CK(cudaMemcpyAsync(&V, P.voff.p + nev, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        TSKB_CK(cudaStreamSynchronize(s));
    }
    P.V = V;
    if ((uint64_t) V + nev + N >= 0xffffffffull) {
        throw (int) TSKB_ERR_UNSUPPORTED;  // 32-bit piece indexes; shard the genome instead
    }
    DevArray<int32_t> vis_node;
    DevArray<double> vis_bl;
    DevArray<uint32_t> vis_ev;
    vis_node.alloc(V); vis_bl.alloc(V); vis_ev.alloc(V);
    if (nev && V) {
        k_chain_fill<<<grid_for(nev, TB), TB, 0, s>>>(ev_parent.p, P.ev_pos.p, nev, a, P.coff.p,
            P.csr_left.p, P.csr_right.p, P.csr_parent.p, P.time.p, P.voff.p, vis_node.p,
            vis_bl.p, vis_ev.p);
        TSKB_CK_LAUNCH();
    }
    ev_parent.release();

    mark("chains");
    // ---- breakpoint index of every event (distinct event positions)
    DevArray<uint32_t> ev_bp;
    ev_bp.alloc(nev + 1);
    {
        DevArray<uint32_t> flag;
        flag.alloc(nev + 1);
        TSKB_CK(cudaMemsetAsync(flag.p, 0, (nev + 1) * sizeof(uint32_t), s));
        if (nev) {
            k_bp_flags<<<grid_for(nev, TB), TB, 0, s>>>(P.ev_pos.p, nev, flag.p);
            TSKB_CK_LAUNCH();
        }
        size_t bytes = 0;
        TSKB_CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, flag.p, ev_bp.p, nev + 1, s));
        TSKB_CK(cub::DeviceScan::ExclusiveSum(tmp.need(bytes), bytes, flag.p, ev_bp.p, nev + 1, s));
        uint32_t T = 0;
        TSKB_CK(cudaMemcpyAsync(&T, ev_bp.p + nev, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        TSKB_CK(cudaStreamSynchronize(s));
        P.T = T;
        P.bp_pos.alloc((size_t) T + 1);
        k_bp_pos<<<grid_for(nev + 1, TB), TB, 0, s>>>(P.ev_pos.p, ev_bp.p, nev, T, b, P.bp_pos.p);
        TSKB_CK_LAUNCH();
    }

    mark("breakpoints");
    // ---- dependency levels: level[parent] > level[child] over every edge
    P.level.alloc(N);
    TSKB_CK(cudaMemsetAsync(P.level.p, 0, N * sizeof(uint32_t), s));
    {
        DevArray<int> changed;
        changed.alloc(1);
        int h_changed = E > 0;
        int rounds = 0;
        while (h_changed) {
            TSKB_CK(cudaMemsetAsync(changed.p, 0, sizeof(int), s));
            for (int r = 0; r < 8; r++) {
                k_relax_levels<<<grid_for(E, TB), TB, 0, s>>>(ep.p, ec.p, E, P.level.p, changed.p);
            }
            TSKB_CK_LAUNCH();
            TSKB_CK(cudaMemcpyAsync(&h_changed, changed.p, sizeof(int), cudaMemcpyDeviceToHost, s));
            TSKB_CK(cudaStreamSynchronize(s));
            if (++rounds > (1 << 20)) break;  // cyclic (invalid) input
        }
    }
    uint32_t max_level = 0;
    {
        DevArray<uint32_t> mx;
        mx.alloc(1);
        size_t bytes = 0;
        if (N) {
            TSKB_CK(cub::DeviceReduce::Max(nullptr, bytes, P.level.p, mx.p, N, s));
            TSKB_CK(cub::DeviceReduce::Max(tmp.need(bytes), bytes, P.level.p, mx.p, N, s));
            TSKB_CK(cudaMemcpyAsync(&max_level, mx.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            TSKB_CK(cudaStreamSynchronize(s));
        }
    }
    P.nlevels = max_level + 1;

    mark("node levels");
    // ---- node rank: nodes sorted by (level, id)
    DevArray<uint32_t> rank, lvl_rank_off;
    rank.alloc(N);
    P.rank_node.alloc(N);
    lvl_rank_off.alloc(P.nlevels + 1);
    {
        DevArray<uint32_t> ids, lvl_sorted;
        ids.alloc(N); lvl_sorted.alloc(N);
        if (N) {
            k_iota<<<grid_for(N, TB), TB, 0, s>>>(ids.p, N);
            TSKB_CK_LAUNCH();
            sort_pairs(tmp, P.level.p, lvl_sorted.p, ids.p, (uint32_t *) P.rank_node.p, N,
                (int) std::max(1u, ceil_log2(P.nlevels + 1)), s);
            k_scatter_rank<<<grid_for(N, TB), TB, 0, s>>>((uint32_t *) P.rank_node.p, N, rank.p);
            TSKB_CK_LAUNCH();
        }
        k_offsets<<<grid_for(P.nlevels + 1, TB), TB, 0, s>>>(lvl_sorted.p, N, P.nlevels + 1,
            lvl_rank_off.p);
        TSKB_CK_LAUNCH();
        TSKB_CK(cudaStreamSynchronize(s));
    }

    mark("node rank");
    // ---- node-major order of the entries (CHILD entries + visits), pieces
    const uint32_t Ve = V + nev;
    DevArray<uint32_t> sorted_e, sorted_key, em_ev, noff, endflag, endscan, inv, poff;
    sorted_e.alloc(Ve); sorted_key.alloc(Ve); em_ev.alloc(Ve); noff.alloc(N + 1);
    endflag.alloc(Ve + 1); endscan.alloc(Ve + 1); inv.alloc(Ve); poff.alloc(N + 1);
    {
        DevArray<uint32_t> kin, vin;
        kin.alloc(Ve); vin.alloc(Ve);
        if (nev) {
            k_entry_keys_child<<<grid_for(nev, TB), TB, 0, s>>>(P.ev_child.p, P.voff.p, rank.p,
                nev, kin.p, em_ev.p);
            TSKB_CK_LAUNCH();
        }
        if (V) {
            k_entry_keys_visit<<<grid_for(V, TB), TB, 0, s>>>(vis_node.p, vis_ev.p, rank.p, V,
                kin.p, em_ev.p);
            TSKB_CK_LAUNCH();
        }
        if (Ve) {
            k_iota<<<grid_for(Ve, TB), TB, 0, s>>>(vin.p, Ve);
            TSKB_CK_LAUNCH();
            // stable: entries of one node stay in event order
            sort_pairs(tmp, kin.p, sorted_key.p, vin.p, sorted_e.p, Ve,
                (int) std::max(1u, ceil_log2(N)), s);
        }
        k_offsets<<<grid_for(N + 1, TB), TB, 0, s>>>(sorted_key.p, Ve, N + 1, noff.p);
        TSKB_CK_LAUNCH();
        TSKB_CK(cudaStreamSynchronize(s));
    }
    vis_node.release(); vis_ev.release();
    uint32_t ends_total = 0;
    {
        TSKB_CK(cudaMemsetAsync(endflag.p, 0, (Ve + 1) * sizeof(uint32_t), s));
        if (Ve) {
            k_entry_ends<<<grid_for(Ve, TB), TB, 0, s>>>(sorted_e.p, sorted_key.p, em_ev.p,
                ev_bp.p, Ve, inv.p, endflag.p);
            TSKB_CK_LAUNCH();
        }
        size_t bytes = 0;
        TSKB_CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, endflag.p, endscan.p, Ve + 1, s));
        TSKB_CK(cub::DeviceScan::ExclusiveSum(tmp.need(bytes), bytes, endflag.p, endscan.p, Ve + 1, s));
        TSKB_CK(cudaMemcpyAsync(&ends_total, endscan.p + Ve, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        TSKB_CK(cudaStreamSynchronize(s));
    }
    P.P = ends_total + N;
    if ((uint64_t) ends_total + N >= 0xfffffff0ull) {
        throw (int) TSKB_ERR_UNSUPPORTED;  // 32-bit piece indexes; shard the genome instead
    }
    // ---- pieces
    // pieces are streamed in whole tiles of 1024 by the summary kernel: pad with INIT markers
    const size_t P_pad = ((size_t) P.P + 1023) / 1024 * 1024;
    // node-major construction arrays (dropped once the processing order is laid out)
    DevArray<double> pc_x, pc_bl;
    pc_x.alloc(P_pad); pc_bl.alloc(P_pad);
    TSKB_CK(cudaMemsetAsync(pc_bl.p, 0, P_pad * sizeof(double), s));
    k_fill_f64<<<grid_for(P_pad - P.P + 1, TB), TB, 0, s>>>(pc_x.p + P.P, P_pad - P.P, -1.0);
    TSKB_CK_LAUNCH();
    DevArray<uint32_t> piece_rank;
    piece_rank.alloc(P.P);
    if (Ve) {
        k_piece_fill<<<grid_for(Ve, TB), TB, 0, s>>>(sorted_e.p, sorted_key.p, em_ev.p, endflag.p,
            endscan.p, P.voff.p, P.ev_sign.p, ev_sbl.p, P.ev_pos.p, vis_bl.p, Ve, pc_x.p,
            pc_bl.p, piece_rank.p);
        TSKB_CK_LAUNCH();
    }
    DevArray<uint8_t> piece_upd;
    piece_upd.alloc((size_t) P.P + 1);
    TSKB_CK(cudaMemsetAsync(piece_upd.p, 0, (size_t) P.P + 1, s));
    if (Ve) {
        k_piece_updates<<<grid_for(Ve, TB), TB, 0, s>>>(sorted_e.p, sorted_key.p, em_ev.p, endscan.p,
            P.voff.p, P.ev_sign.p, Ve, piece_upd.p);
        TSKB_CK_LAUNCH();
    }
    k_piece_init<<<grid_for(N + 1, TB), TB, 0, s>>>(noff.p, endscan.p, Ve, ends_total, N, pc_x.p,
        pc_bl.p, piece_rank.p, poff.p);
    TSKB_CK_LAUNCH();
    TSKB_CK(cudaStreamSynchronize(s));
    ev_sbl.release(); vis_bl.release(); inv.release(); em_ev.release(); sorted_e.release();
    ev_bp.release(); endflag.release(); sorted_key.release(); endscan.release(); noff.release();

    mark("entries + pieces");
    // ---- parent-major edge CSR (children of a node at a position; also used by the decode)
    P.pm_off.alloc(N + 1); P.pm_left.alloc(E); P.pm_right.alloc(E); P.pm_pmax.alloc(E);
    P.pm_child.alloc(E);
    {
        DevArray<uint64_t> k64, k64o;
        DevArray<uint32_t> v, vo, pk, pko, perm;
        DevArray<int32_t> child_of;
        k64.alloc(E); k64o.alloc(E); v.alloc(E); vo.alloc(E); pk.alloc(E); pko.alloc(E);
        perm.alloc(E); child_of.alloc(E);
        if (E) {
            k_pm_keys<<<grid_for(E, TB), TB, 0, s>>>(P.coff.p, N, P.csr_left.p, E, k64.p, v.p,
                child_of.p);
            TSKB_CK_LAUNCH();
            sort_pairs(tmp, k64.p, k64o.p, v.p, vo.p, E, 64, s);
            k_gather_parent<<<grid_for(E, TB), TB, 0, s>>>(vo.p, P.csr_parent.p, E, pk.p);
            TSKB_CK_LAUNCH();
            sort_pairs(tmp, pk.p, pko.p, vo.p, perm.p, E, (int) std::max(1u, ceil_log2(N)), s);
            k_pm_gather<<<grid_for(E, TB), TB, 0, s>>>(perm.p, E, P.csr_left.p, P.csr_right.p,
                child_of.p, P.pm_left.p, P.pm_right.p, P.pm_child.p);
            TSKB_CK_LAUNCH();
            size_t bytes = 0;
            TSKB_CK(cub::DeviceScan::InclusiveScanByKey(nullptr, bytes, pko.p, P.pm_right.p,
                P.pm_pmax.p, MaxOp(), E, ::cuda::std::equal_to<>(), s));
            TSKB_CK(cub::DeviceScan::InclusiveScanByKey(tmp.need(bytes), bytes, pko.p, P.pm_right.p,
                P.pm_pmax.p, MaxOp(), E, ::cuda::std::equal_to<>(), s));
        }
        k_offsets<<<grid_for(N + 1, TB), TB, 0, s>>>(pko.p, E, N + 1, P.pm_off.p);
        TSKB_CK_LAUNCH();
        TSKB_CK(cudaStreamSynchronize(s));
    }

    mark("parent-major CSR");
    // ---- mutations: the (node-major) piece holding state[mutation.node] at the site; needed pieces
    DevArray<uint8_t> needed;
    needed.alloc((size_t) P.P + 1);
    P.all_pieces = (options & TSKB_INIT_NODE_MODE) != 0;
    k_needed_branch<<<grid_for(P.P, TB), TB, 0, s>>>(P.P, pc_x.p, pc_bl.p, P.all_pieces ? 1 : 0, needed.p);
    TSKB_CK_LAUNCH();
    P.mut_src.alloc(P.Mu);
    if (P.Mu) {
        const uint32_t Mu = (uint32_t) P.Mu;
        k_mut_src<<<grid_for(Mu, TB), TB, 0, s>>>(d_msite.p, P.mut_node.p, Mu, P.site_pos.p, rank.p,
            poff.p, pc_x.p, P.mut_src.p);
        k_needed_mutation<<<grid_for(Mu, TB), TB, 0, s>>>(Mu, P.mut_src.p, pc_x.p, needed.p);
        TSKB_CK_LAUNCH();
        TSKB_CK(cudaStreamSynchronize(s));
    }

    mark("mutation sources");
    // ---- references of every piece, heights, processing order
    DevArray<uint32_t> perm;  // node-major piece -> state slot
    {
        const uint32_t Pn = P.P;
        DevArray<uint32_t> is_sample, cnt, ch_off, refs, height;
        is_sample.alloc(N); cnt.alloc((size_t) Pn + 1); ch_off.alloc((size_t) Pn + 1);
        height.alloc(Pn);
        TSKB_CK(cudaMemsetAsync(is_sample.p, 0, N * sizeof(uint32_t), s));
        TSKB_CK(cudaMemsetAsync(cnt.p, 0, ((size_t) Pn + 1) * sizeof(uint32_t), s));
        TSKB_CK(cudaMemsetAsync(height.p, 0, (size_t) Pn * sizeof(uint32_t), s));
        if (P.num_samples) {
            k_flag_samples<<<grid_for(P.num_samples, TB), TB, 0, s>>>(P.d_samples.p, P.num_samples,
                is_sample.p);
            TSKB_CK_LAUNCH();
        }
        k_piece_children<false><<<grid_for(Pn, TB), TB, 0, s>>>(Pn, pc_x.p, piece_rank.p,
            P.rank_node.p, rank.p, is_sample.p, poff.p, P.pm_off.p, P.pm_left.p, P.pm_right.p,
            P.pm_pmax.p, P.pm_child.p, cnt.p, nullptr, nullptr);
        TSKB_CK_LAUNCH();
        if (sum_u64(tmp, cnt.p, (size_t) Pn + 1, s) >= 0xfffffff0ull) {
            throw (int) TSKB_ERR_UNSUPPORTED;  // 32-bit reference offsets; shard the genome instead
        }
        size_t bytes = 0;
        TSKB_CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt.p, ch_off.p, (size_t) Pn + 1, s));
        TSKB_CK(cub::DeviceScan::ExclusiveSum(tmp.need(bytes), bytes, cnt.p, ch_off.p, (size_t) Pn + 1, s));
        uint32_t nrefs = 0;
        TSKB_CK(cudaMemcpyAsync(&nrefs, ch_off.p + Pn, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        TSKB_CK(cudaStreamSynchronize(s));
        refs.alloc((size_t) nrefs + 1);
        k_piece_children<true><<<grid_for(Pn, TB), TB, 0, s>>>(Pn, pc_x.p, piece_rank.p,
            P.rank_node.p, rank.p, is_sample.p, poff.p, P.pm_off.p, P.pm_left.p, P.pm_right.p,
            P.pm_pmax.p, P.pm_child.p, cnt.p, ch_off.p, refs.p);
        TSKB_CK_LAUNCH();
        {
            DevArray<int> changed;
            changed.alloc(1);
            int h_changed = 1, rounds = 0;
            while (h_changed) {
                TSKB_CK(cudaMemsetAsync(changed.p, 0, sizeof(int), s));
                for (int q = 0; q < 4; q++) {
                    k_relax_height<<<grid_for(Pn, TB), TB, 0, s>>>(Pn, ch_off.p, refs.p, pc_x.p,
                        height.p, changed.p);
                }
                TSKB_CK_LAUNCH();
                TSKB_CK(cudaMemcpyAsync(&h_changed, changed.p, sizeof(int), cudaMemcpyDeviceToHost, s));
                TSKB_CK(cudaStreamSynchronize(s));
                if (++rounds > (1 << 20)) throw (int) TSKB_ERR_BAD_PARAM_VALUE;  // cyclic input
            }
        }
        const char *order_env = getenv("TSKB_ORDER");
        const bool order_x = order_env != nullptr && order_env[0] == 'x';
        const bool order_level = order_env != nullptr && order_env[0] == 'l';
        if (order_level && Pn) {
            // group by the node's static level instead of the piece's height: more dependent steps,
            // but every node's pieces stay contiguous
            k_level_as_height<<<grid_for(Pn, TB), TB, 0, s>>>(Pn, piece_rank.p, P.rank_node.p, P.level.p,
                height.p);
            TSKB_CK_LAUNCH();
        }
        uint32_t max_h = 0;
        {
            DevArray<uint32_t> mx;
            mx.alloc(1);
            if (Pn) {
                TSKB_CK(cub::DeviceReduce::Max(nullptr, bytes, height.p, mx.p, Pn, s));
                TSKB_CK(cub::DeviceReduce::Max(tmp.need(bytes), bytes, height.p, mx.p, Pn, s));
                TSKB_CK(cudaMemcpyAsync(&max_h, mx.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
                TSKB_CK(cudaStreamSynchronize(s));
            }
        }
        P.nheights = max_h + 1;
        // sort pieces by height (stable: node-major piece order within a height, or by
        // position with TSKB_ORDER=x); INIT pieces last
        DevArray<uint32_t> kin, kout, vin, vout, lvl_begin;
        kin.alloc(Pn); kout.alloc(Pn); vin.alloc(Pn); vout.alloc(Pn); lvl_begin.alloc(P.nheights + 2);
        if (Pn) {
            if (order_x) {
                DevArray<uint64_t> k64, k64o;
                DevArray<uint32_t> v0;
                k64.alloc(Pn); k64o.alloc(Pn); v0.alloc(Pn);
                k_x_keys<<<grid_for(Pn, TB), TB, 0, s>>>(Pn, pc_x.p, k64.p, v0.p);
                TSKB_CK_LAUNCH();
                sort_pairs(tmp, k64.p, k64o.p, v0.p, vin.p, Pn, 64, s);
                k_height_keys_of<<<grid_for(Pn, TB), TB, 0, s>>>(Pn, vin.p, needed.p, height.p, kin.p,
                    P.nheights);
                TSKB_CK_LAUNCH();
                TSKB_CK(cudaStreamSynchronize(s));
            } else {
                k_height_keys<<<grid_for(Pn, TB), TB, 0, s>>>(Pn, needed.p, height.p, kin.p, vin.p,
                    P.nheights);
                TSKB_CK_LAUNCH();
            }
            sort_pairs(tmp, kin.p, kout.p, vin.p, vout.p, Pn,
                (int) std::max(1u, ceil_log2(P.nheights + 2)), s);
        }
        k_offsets<<<grid_for(P.nheights + 2, TB), TB, 0, s>>>(kout.p, Pn, P.nheights + 2, lvl_begin.p);
        TSKB_CK_LAUNCH();
        std::vector<uint32_t> h_begin = lvl_begin.download(s);  // sorted offset of each height
        const uint32_t nreal = h_begin[P.nheights];             // INIT pieces follow
        std::vector<uint32_t> h_padded(P.nheights + 1), h_dep;
        P.level_begin.assign(P.nheights + 1, 0);
        uint64_t padded = 0;
        for (uint32_t h = 0; h < P.nheights; h++) {
            h_padded[h] = (uint32_t) padded;
            P.level_begin[h] = (uint32_t) padded;
            const uint32_t dep = (uint32_t) h_dep.size();
            const uint32_t len = h_begin[h + 1] - h_begin[h];
            for (uint32_t u = 0; u < len; u += PROP_TILE) h_dep.push_back(dep);
            padded = (uint64_t) h_dep.size() * PROP_TILE;
            if (padded + P.num_samples + 1 >= 0x7fffffffull) throw (int) TSKB_ERR_UNSUPPORTED;
        }
        h_padded[P.nheights] = (uint32_t) padded;
        P.level_begin[P.nheights] = (uint32_t) padded;
        P.npp = (uint32_t) padded;
        P.ntiles = (uint32_t) h_dep.size();
        P.tile_dep.upload(h_dep.data(), h_dep.size(), s);
        DevArray<uint32_t> d_padded, q_cnt;
        d_padded.upload(h_padded.data(), h_padded.size(), s);
        P.d_sample_index.upload(P.sample_index_map.data(), N, s);
        P.q_bp0.alloc(P.npp); P.q_bp1.alloc(P.npp); P.q_bl.alloc(P.npp); P.q_eff.alloc(P.npp);
        P.q_off.alloc((size_t) P.npp + 1); q_cnt.alloc((size_t) P.npp + 1);
        perm.alloc((size_t) Pn + 1);
        // default: the zero slot (pieces that are not computed are never referenced)
        k_fill_u32<<<grid_for((size_t) Pn + 1, TB), TB, 0, s>>>(perm.p, (size_t) Pn + 1,
            P.npp + P.num_samples);
        TSKB_CK_LAUNCH();
        TSKB_CK(cudaMemsetAsync(q_cnt.p, 0, ((size_t) P.npp + 1) * sizeof(uint32_t), s));
        if (P.npp) {
            TSKB_CK(cudaMemsetAsync(P.q_bp0.p, 0, (size_t) P.npp * sizeof(uint32_t), s));
            TSKB_CK(cudaMemsetAsync(P.q_bl.p, 0, (size_t) P.npp * sizeof(double), s));
            // padding entries are marked by their end breakpoint
            k_fill_u32<<<grid_for(P.npp, TB), TB, 0, s>>>(P.q_bp1.p, P.npp, NO_PIECE);
            TSKB_CK_LAUNCH();
        }
        if (nreal) {
            k_order_fill<<<grid_for(nreal, TB), TB, 0, s>>>(vout.p, kout.p, nreal, lvl_begin.p,
                d_padded.p, cnt.p, pc_x.p, pc_bl.p, Pn, P.bp_pos.p, P.T, piece_upd.p, P.q_bp0.p, P.q_bp1.p,
                P.q_eff.p, P.q_bl.p, q_cnt.p, perm.p);
            TSKB_CK_LAUNCH();
        }
        if (N) {
            k_init_perm<<<grid_for(N, TB), TB, 0, s>>>(N, poff.p, P.rank_node.p,
                P.d_sample_index.p, P.npp, P.num_samples, perm.p);
            TSKB_CK_LAUNCH();
        }
        TSKB_CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, q_cnt.p, P.q_off.p, (size_t) P.npp + 1, s));
        TSKB_CK(cub::DeviceScan::ExclusiveSum(tmp.need(bytes), bytes, q_cnt.p, P.q_off.p,
            (size_t) P.npp + 1, s));
        P.nrefs = nrefs;
        P.refs.alloc((size_t) nrefs + 16);
        if (nreal) {
            k_refs_reorder<<<grid_for(nreal, TB), TB, 0, s>>>(nreal, vout.p, perm.p, P.q_off.p,
                ch_off.p, refs.p, P.refs.p);
            TSKB_CK_LAUNCH();
        }
        if (P.all_pieces) {
            P.q_node.alloc(P.npp);
            P.node_first_bp.alloc(N);
            if (P.npp) {
                TSKB_CK(cudaMemsetAsync(P.q_node.p, 0xff, (size_t) P.npp * sizeof(int32_t), s));
                k_piece_nodes<<<grid_for(Pn, TB), TB, 0, s>>>(Pn, pc_x.p, piece_rank.p, P.rank_node.p, perm.p,
                    P.npp, P.q_node.p);
                TSKB_CK_LAUNCH();
            }
            if (N) {
                k_node_first_bp<<<grid_for(N, TB), TB, 0, s>>>(N, poff.p, P.rank_node.p, pc_x.p, P.bp_pos.p,
                    P.T, P.node_first_bp.p);
                TSKB_CK_LAUNCH();
            }
        }
        TSKB_CK(cudaStreamSynchronize(s));
    }

    mark("references + heights + order");
    // ---- sites and mutations: allele strings -> small integer codes (k_allele_codes)
    {
        const uint32_t S = (uint32_t) P.S, Mu = (uint32_t) P.Mu;
        P.h_site_pos.assign(t->site_position, t->site_position + S);
        P.site_lo = (uint32_t) (std::lower_bound(P.h_site_pos.begin(), P.h_site_pos.end(), a) - P.h_site_pos.begin());
        P.site_hi = (uint32_t) (std::lower_bound(P.h_site_pos.begin(), P.h_site_pos.end(), b) - P.h_site_pos.begin());
        P.site_moff.alloc((size_t) S + 1);
        P.site_aoff.alloc((size_t) S + 1);
        P.mut_allele.alloc(Mu);
        P.mut_alt.alloc(Mu);
        // mutation CSR by site: mutation_site is sorted (checked above)
        k_offsets<<<grid_for((size_t) S + 1, TB), TB, 0, s>>>((const uint32_t *) d_msite.p, Mu, S + 1, P.site_moff.p);
        TSKB_CK_LAUNCH();
        P.total_alleles = 0;
        if (S) {
            DevArray<char> d_anc, d_der;
            DevArray<uint64_t> d_anc_off, d_der_off;
            DevArray<uint32_t> nall, red;
            DevArray<int32_t> d_mpar;
            DevArray<int> ovf;
            d_anc.upload(t->site_ancestral_state, t->site_ancestral_state_offset[S], s);
            d_anc_off.upload(t->site_ancestral_state_offset, (size_t) S + 1, s);
            if (Mu) {
                d_der.upload(t->mutation_derived_state, t->mutation_derived_state_offset[Mu], s);
                d_der_off.upload(t->mutation_derived_state_offset, (size_t) Mu + 1, s);
                if (t->mutation_parent != nullptr) d_mpar.upload(t->mutation_parent, Mu, s);
            }
            nall.alloc((size_t) S + 1); red.alloc(2); ovf.alloc(1);
            TSKB_CK(cudaMemsetAsync(nall.p, 0, ((size_t) S + 1) * sizeof(uint32_t), s));
            TSKB_CK(cudaMemsetAsync(ovf.p, 0, sizeof(int), s));
            k_allele_codes<<<grid_for(S, 128), 128, 0, s>>>(S, P.site_moff.p, d_anc.p, d_anc_off.p, d_der.p,
                d_der_off.p, d_mpar.p, P.mut_allele.p, P.mut_alt.p, nall.p, ovf.p);
            TSKB_CK_LAUNCH();
            size_t bytes = 0;
            TSKB_CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, nall.p, P.site_aoff.p, (size_t) S + 1, s));
            TSKB_CK(cub::DeviceScan::ExclusiveSum(tmp.need(bytes), bytes, nall.p, P.site_aoff.p, (size_t) S + 1, s));
            TSKB_CK(cub::DeviceReduce::Max(nullptr, bytes, nall.p, red.p, S, s));
            TSKB_CK(cub::DeviceReduce::Max(tmp.need(bytes), bytes, nall.p, red.p, S, s));
            // mutations per site: differences of the CSR offsets
            DevArray<uint32_t> nm;
            nm.alloc(S);
            k_adjacent_diff<<<grid_for(S, TB), TB, 0, s>>>(P.site_moff.p, S, nm.p);
            TSKB_CK_LAUNCH();
            TSKB_CK(cub::DeviceReduce::Max(nullptr, bytes, nm.p, red.p + 1, S, s));
            TSKB_CK(cub::DeviceReduce::Max(tmp.need(bytes), bytes, nm.p, red.p + 1, S, s));
            uint32_t h_red[2] = { 1, 0 }, h_total = 0;
            int h_ovf = 0;
            TSKB_CK(cudaMemcpyAsync(h_red, red.p, sizeof(h_red), cudaMemcpyDeviceToHost, s));
            TSKB_CK(cudaMemcpyAsync(&h_total, P.site_aoff.p + S, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            TSKB_CK(cudaMemcpyAsync(&h_ovf, ovf.p, sizeof(int), cudaMemcpyDeviceToHost, s));
            TSKB_CK(cudaStreamSynchronize(s));
            if (h_ovf) throw (int) TSKB_ERR_UNSUPPORTED;
            P.max_alleles_per_site = std::max<uint32_t>(1, h_red[0]);
            P.max_muts_per_site = h_red[1];
            P.total_alleles = h_total;
        } else {
            TSKB_CK(cudaMemsetAsync(P.site_aoff.p, 0, sizeof(uint32_t), s));
        }
        if (Mu) {
            k_translate<<<grid_for(Mu, TB), TB, 0, s>>>(P.mut_src.p, Mu, perm.p);
            TSKB_CK_LAUNCH();
        }
        TSKB_CK(cudaStreamSynchronize(s));
    }
    TSKB_CK(cudaStreamSynchronize(s));

    mark("allele codes");
    P.stats.num_events = nev;
    P.stats.num_visits = V;
    P.stats.num_levels = P.nheights;
    P.stats.device_bytes = P.device_bytes();
    P.stats.stage_ms = std::chrono::duration<double, std::milli>(
        std::chrono::steady_clock::now() - t_start).count();
    return plan.release();
}

}  // namespace tskb
