// matrix.cu -- genotype decode and the site-mode divergence matrix (a9, a10, a12).
//
// Decode (tsk_variant_decode over all sites, c/tskit/genotypes.c:473-594): every genotype
// starts at the ancestral allele 0 (genotypes.c:550-552), isolated samples are marked -1 unless
// TSK_ISOLATED_NOT_MISSING (genotypes.c:413-455, 554-559), then the site's mutations are applied
// in table order, each overwriting the listed samples in the subtree below its node
// (genotypes.c:355-411).  On the device the subtree walk is a breadth-first expansion of
// (site, node) items over a parent-major edge CSR -- output-sensitive: the work is the number
// of non-ancestral genotypes, not sites x samples -- one round per mutation rank so that later
// mutations of a site overwrite earlier ones exactly as the reference's loop does.
//
// Divergence matrix, site mode (trees.c:8684-8826, 8876-8899): for every site of the window and
// every pair of samples carrying different alleles the pair's entry grows by one.  With the
// genotypes sample-major, "same allele" counts are  sum over alleles a of (G == a)(G == a)^T, an
// int8 x int8 -> int32 product (exact), and different = sites - same.  Set aggregation,
// count-normalisation and span-normalisation follow in fp64.
#include <cuda.h>
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "plan.cuh"

namespace tskb {

namespace {

constexpr int TB = 256;

__global__ void k_fill_i32(int32_t *out, size_t n, int32_t v) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}

__global__ void k_set_cols(const int32_t *samples, uint32_t n, int32_t *col) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) col[samples[i]] = (int32_t) i;
}

// ---- isolated samples: no edge above and no edge below at the site's position
__global__ void k_mark_isolated(const int32_t *samples, uint32_t n, const uint32_t *coff,
    const double *csr_left, const double *csr_right, const uint32_t *pm_off, const double *pm_left,
    const double *pm_right, const double *site_pos, uint32_t s0, uint32_t s1, double L, int8_t *G,
    size_t stride_site, size_t stride_sample) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t u = samples[i];
    // merge the two interval lists (both sorted by left); reach = right end of the coverage
    uint32_t a = coff[u], a1 = coff[u + 1], b = pm_off[u], b1 = pm_off[u + 1];
    double reach = 0.0;
    while (true) {
        double nl, nr;
        bool ha = a < a1, hb = b < b1;
        if (!ha && !hb) {
            nl = L;
            nr = L;
        } else if (ha && (!hb || csr_left[a] <= pm_left[b])) {
            nl = csr_left[a];
            nr = csr_right[a];
            a++;
        } else {
            nl = pm_left[b];
            nr = pm_right[b];
            b++;
        }
        if (nl > reach) {  // gap [reach, nl): isolated
            uint32_t lo = lower_bound_dev(site_pos + s0, s1 - s0, reach) + s0;
            uint32_t hi = lower_bound_dev(site_pos + s0, s1 - s0, nl) + s0;
            for (uint32_t s = lo; s < hi; s++) G[(size_t) (s - s0) * stride_site + (size_t) i * stride_sample] = -1;
        }
        if (nr > reach) reach = nr;
        if (!ha && !hb) break;
    }
}

// ---- breadth-first expansion of (site, node) items
__global__ void k_frontier_init(uint32_t r, uint32_t s0, uint32_t s1, const uint32_t *site_moff,
    const int32_t *mut_node, uint2 *items, uint32_t *count) {
    uint32_t s = s0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= s1) return;
    uint32_t m0 = site_moff[s], m1 = site_moff[s + 1];
    if (m1 - m0 > r) {
        uint32_t idx = atomicAdd(count, 1u);
        items[idx] = make_uint2(s, (uint32_t) mut_node[m0 + r]);
    }
}

__global__ void k_expand(const uint2 *items, uint32_t nitems, uint32_t r, uint32_t s0,
    const double *site_pos, const uint32_t *site_moff, const uint16_t *mut_allele,
    const int32_t *col, const uint32_t *pm_off, const double *pm_left, const double *pm_right,
    const double *pm_pmax, const int32_t *pm_child, int8_t *G, size_t stride_site,
    size_t stride_sample, const uint32_t *site_col, uint2 *next, uint32_t cap, uint32_t *next_count,
    int *overflow) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nitems) return;
    const uint2 it = items[t];
    const uint32_t s = it.x, u = it.y;
    const double x = site_pos[s];
    int32_t c = col[u];
    if (c >= 0) {
        const size_t column = site_col != nullptr ? site_col[s - s0] : s - s0;  // padded layouts
        G[column * stride_site + (size_t) c * stride_sample] = (int8_t) mut_allele[site_moff[s] + r];
    }
    uint32_t lo = pm_off[u], hi = pm_off[u + 1];
    uint32_t k = upper_bound_dev(pm_left + lo, hi - lo, x);  // edges with left <= x
    for (uint32_t j = lo + k; j-- > lo;) {
        if (!(pm_pmax[j] > x)) break;  // nothing further left reaches x
        if (pm_right[j] > x) {
            uint32_t idx = atomicAdd(next_count, 1u);
            if (idx < cap) {
                next[idx] = make_uint2(s, (uint32_t) pm_child[j]);
            } else {
                *overflow = 1;
            }
        }
    }
}

// ---- same-allele counts: C[i][j] += sum over sites k in [k_lo, k_hi) of [X[i][k] == a][X[j][k] == a]
// X sample-major int8 [n x ld].  128 x 128 x 64 CTA tiles, 8 warps of 64 x 32, legacy int8
// tensor path (mma.sync m16n8k32, exact int32 accumulation); only blocks with bi <= bj.
constexpr int GM = 128, GK = 64, GLD = GK + 16;

__device__ __forceinline__ void mma_s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
    uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void load_onehot_tile(const int8_t *X, size_t ld, uint32_t n, uint32_t row0,
    uint32_t k0, uint32_t k_lo, uint32_t k_hi, uint32_t a_rep, unsigned char (*dst)[GLD]) {
    // 128 rows x 64 bytes = 512 chunks of 16 bytes, 2 per thread
#pragma unroll
    for (int c = 0; c < 2; c++) {
        uint32_t chunk = threadIdx.x + c * TB;
        uint32_t row = chunk >> 2, kc = (chunk & 3) * 16;
        uint4 v = make_uint4(0, 0, 0, 0);
        uint32_t k = k0 + kc;
        if (row0 + row < n && k < k_hi && k + 16 > k_lo) {
            uint4 raw = *reinterpret_cast<const uint4 *>(X + (size_t) (row0 + row) * ld + k);
            uint32_t wds[4] = { raw.x, raw.y, raw.z, raw.w };
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t eq = __vcmpeq4(wds[q], a_rep) & 0x01010101u;
                uint32_t kb = k + 4 * q;
                if (kb < k_lo || kb + 4 > k_hi) {  // partial word at the range ends
                    uint32_t m = 0;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        if (kb + b >= k_lo && kb + b < k_hi) m |= 0xffu << (8 * b);
                    }
                    eq &= m;
                }
                wds[q] = eq;
            }
            v = make_uint4(wds[0], wds[1], wds[2], wds[3]);
        }
        *reinterpret_cast<uint4 *>(&dst[row][kc]) = v;
    }
}

__global__ void __launch_bounds__(TB) k_same_gemm(const int8_t *X, size_t ld, uint32_t n, uint32_t k_lo,
    uint32_t k_hi, int allele, int32_t *C) {
    __shared__ __align__(16) unsigned char As[GM][GLD];
    __shared__ __align__(16) unsigned char Bs[GM][GLD];
    const uint32_t bi = blockIdx.y, bj = blockIdx.x;
    if (bi > bj) return;
    const uint32_t i0 = bi * GM, j0 = bj * GM;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t wm = warp >> 2, wn = warp & 3, g = lane >> 2, tig = lane & 3;
    const uint32_t a_rep = 0x01010101u * (uint32_t) (allele & 0xff);
    int acc[4][4][4];
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++)
#pragma unroll
            for (int q = 0; q < 4; q++) acc[mi][ni][q] = 0;
    for (uint32_t k0 = k_lo & ~15u; k0 < k_hi; k0 += GK) {
        load_onehot_tile(X, ld, n, i0, k0, k_lo, k_hi, a_rep, As);
        load_onehot_tile(X, ld, n, j0, k0, k_lo, k_hi, a_rep, Bs);
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < GK / 32; ks++) {
            uint32_t bf[4][2];
#pragma unroll
            for (int ni = 0; ni < 4; ni++) {
                const unsigned char *p = &Bs[wn * 32 + ni * 8 + g][ks * 32 + tig * 4];
                bf[ni][0] = *reinterpret_cast<const uint32_t *>(p);
                bf[ni][1] = *reinterpret_cast<const uint32_t *>(p + 16);
            }
#pragma unroll
            for (int mi = 0; mi < 4; mi++) {
                const unsigned char *p = &As[wm * 64 + mi * 16 + g][ks * 32 + tig * 4];
                uint32_t a0 = *reinterpret_cast<const uint32_t *>(p);
                uint32_t a1 = *reinterpret_cast<const uint32_t *>(p + 8 * GLD);
                uint32_t a2 = *reinterpret_cast<const uint32_t *>(p + 16);
                uint32_t a3 = *reinterpret_cast<const uint32_t *>(p + 8 * GLD + 16);
#pragma unroll
                for (int ni = 0; ni < 4; ni++) mma_s8(acc[mi][ni], a0, a1, a2, a3, bf[ni][0], bf[ni][1]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            uint32_t r0 = i0 + wm * 64 + mi * 16 + g, c0 = j0 + wn * 32 + ni * 8 + 2 * tig;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t r = r0 + (q >> 1) * 8, c = c0 + (q & 1);
                if (r < n && c < n) C[(size_t) r * n + c] += acc[mi][ni][q];
            }
        }
    }
}

// ---- the same contraction on the 5th-generation tensor cores (tcgen05, sm_100a)
// C[i][j] = sum over alleles a < num_alleles and sites k in [k_lo, k_hi) of [X[i][k] == a][X[j][k] == a].
// One CTA per 128 x 128 block of C (row block <= column block).  All 256 threads turn 128-byte
// row segments of X into int8 one-hot tiles in shared memory, written directly in the K-major
// "interleaved" (no-swizzle) UMMA layout: 8 x 16-byte core matrices, 128 bytes each, at
// (row / 8) * SBO + (k / 16) * LBO.  One elected thread issues tcgen05.mma.kind::i8
// (M = 128, N = 128, K = 32 per instruction, u8 x u8 -> s32, exact) with the accumulator in tensor
// memory (128 lanes x 128 columns) for every allele and k-chunk, and commits to an mbarrier; the
// tile stage is double-buffered so that the next one-hot tiles are built while the tensor core
// works on the current ones.  Epilogue: tcgen05.ld of the accumulator (warp w reads TMEM lanes
// 32 (w % 4) ..., columns 64 (w / 4) ...) and 128-byte row-segment stores.
namespace tc {

constexpr uint32_t TM = 128, TN = 128, BK = 128;          // CTA tile, bytes of K per stage
constexpr uint32_t TILE_BYTES = TM * BK;                   // one operand tile: 16 KB
constexpr uint32_t LBO = 128, SBO = (BK / 16) * 128;       // core-matrix strides (bytes)
constexpr uint32_t STAGES = 2;
constexpr uint32_t TMEM_COLS = 128;
constexpr uint32_t SMEM_BYTES = STAGES * 2 * TILE_BYTES + 1024;  // + alignment slack
constexpr int THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t) __cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) { }
}
// K-major, no swizzle (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp): start address, leading
// (K) and stride (M/N) byte offsets in 16-byte units, descriptor version 1 in bits [46,48)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t) ((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t) ((LBO >> 4) & 0x3fffu) << 16;
    d |= (uint64_t) ((SBO >> 4) & 0x3fffu) << 32;
    d |= (uint64_t) 1 << 46;
    return d;
}
// cute::UMMA::InstrDescriptor for kind::i8: dense, no saturate, D = S32 (2) at [4,6), A/B = unsigned
// 8 bit (0) at [7,10)/[10,13), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t IDESC = (2u << 4) | ((TN >> 3) << 17) | ((TM >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar)) : "memory");
}

// 16 bytes of genotypes -> 16 one-hot bytes for allele a, bytes outside [k_lo, k_hi) cleared
__device__ __forceinline__ uint4 onehot16(uint4 raw, uint32_t a_rep, uint32_t k, uint32_t k_lo, uint32_t k_hi) {
    uint32_t wds[4] = { raw.x, raw.y, raw.z, raw.w };
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t eq = __vcmpeq4(wds[q], a_rep) & 0x01010101u;
        const uint32_t kb = k + 4 * q;
        if (kb < k_lo || kb + 4 > k_hi) {  // partial word at the range ends
            uint32_t m = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                if (kb + b >= k_lo && kb + b < k_hi) m |= 0xffu << (8 * b);
            }
            eq &= m;
        }
        wds[q] = eq;
    }
    return make_uint4(wds[0], wds[1], wds[2], wds[3]);
}

__global__ void __launch_bounds__(THREADS) k_same_umma(const int8_t *__restrict__ X, size_t ld, uint32_t n,
    uint32_t k_lo, uint32_t k_hi, uint32_t num_alleles, int32_t *__restrict__ C) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t s_bar[STAGES];
    __shared__ uint32_t s_tmem;
    const uint32_t bi = blockIdx.y, bj = blockIdx.x;
    if (bi > bj) return;
    const bool diag = bi == bj;
    const uint32_t i0 = bi * TM, j0 = bj * TN;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char *tiles = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);  // 1 KB aligned

    if (tid == 0) {
        for (uint32_t st = 0; st < STAGES; st++) mbar_init(&s_bar[st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&s_tmem)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = s_tmem;

    // this thread's 16-byte chunks of a tile: lanes cover 8 rows x 4 chunks (conflict-free 16-byte
    // shared stores, 64 contiguous bytes per row from global), 4 passes
    const uint32_t r8 = lane & 7, kq = lane >> 3;
    const uint32_t k_begin = k_lo & ~15u;
    const uint32_t nchunks = (k_hi - k_begin + BK - 1) / BK;
    uint32_t it = 0;
    for (uint32_t ch = 0; ch < nchunks; ch++) {
        const uint32_t k0 = k_begin + ch * BK;
        uint4 rawA[4], rawB[4];
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const uint32_t u = warp + 8 * p, g = u >> 1, kc = (u & 1) * 4 + kq;
            const uint32_t row = g * 8 + r8, k = k0 + kc * 16;
            rawA[p] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);  // matches no allele
            rawB[p] = rawA[p];
            if (k < k_hi) {
                if (i0 + row < n) rawA[p] = *reinterpret_cast<const uint4 *>(X + (size_t) (i0 + row) * ld + k);
                if (!diag && j0 + row < n) rawB[p] = *reinterpret_cast<const uint4 *>(X + (size_t) (j0 + row) * ld + k);
            }
        }
        for (uint32_t al = 0; al < num_alleles; al++, it++) {
            const uint32_t st = it % STAGES;
            // the MMAs that read this stage two iterations ago must have finished
            if (it >= STAGES) mbar_wait(&s_bar[st], ((it / STAGES) - 1) & 1);
            unsigned char *tA = tiles + (size_t) st * 2 * TILE_BYTES;
            unsigned char *tB = tA + TILE_BYTES;
            const uint32_t a_rep = 0x01010101u * (al & 0xffu);
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const uint32_t u = warp + 8 * p, g = u >> 1, kc = (u & 1) * 4 + kq;
                const uint32_t off = g * SBO + kc * LBO + r8 * 16;
                const uint32_t k = k0 + kc * 16;
                *reinterpret_cast<uint4 *>(tA + off) = onehot16(rawA[p], a_rep, k, k_lo, k_hi);
                if (!diag) *reinterpret_cast<uint4 *>(tB + off) = onehot16(rawB[p], a_rep, k, k_lo, k_hi);
            }
            // generic-proxy writes -> visible to the tensor core (async proxy)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t aaddr = smem_u32(tA), baddr = smem_u32(diag ? tA : tB);
#pragma unroll
                for (uint32_t ks = 0; ks < BK / 32; ks++) {
                    // K = 32 bytes per instruction = two 16-byte core-matrix columns
                    umma_i8(tmem_acc, umma_desc(aaddr + ks * 2 * LBO), umma_desc(baddr + ks * 2 * LBO),
                        (it > 0 || ks > 0) ? 1u : 0u);
                }
                umma_commit(&s_bar[st]);  // arrives when the MMAs issued so far are complete
            }
        }
    }
    // all MMAs done: the last commit of every stage used
    for (uint32_t back = 0; back < STAGES && back < it; back++) {
        const uint32_t last = it - 1 - back;
        mbar_wait(&s_bar[last % STAGES], (last / STAGES) & 1);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w owns TMEM lanes 32 (w % 4) .. + 31 (its rows) and columns 64 (w / 4) .. + 63
    const uint32_t row = i0 + 32 * (warp & 3) + lane;
#pragma unroll
    for (uint32_t half = 0; half < 2; half++) {
        const uint32_t col0 = 64 * (warp >> 2) + 32 * half;
        uint32_t v[32];
        const uint32_t taddr = tmem_acc + ((32 * (warp & 3)) << 16) + col0;
        if (it > 0) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
            for (int q = 0; q < 32; q++) v[q] = 0;
        }
        if (row < n) {
#pragma unroll
            for (int q = 0; q < 32; q++) {
                const uint32_t col = j0 + col0 + q;
                if (col < n) C[(size_t) row * n + col] = (int32_t) v[q];
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(TMEM_COLS) : "memory");
    }
}

// ---- biallelic sites: C = G G^T straight from the genotype bytes
// With genotypes in {0, 1} the one-hot tile of allele 1 is the genotype tile itself and
//   different[i][j] = C[i][i] + C[j][j] - 2 C[i][j]          (C[i][i] = number of 1s of sample i),
// so the contraction needs neither the compare nor allele 0.  One CTA per 256 x 256 block of C (row
// block <= column block): four 128 x 128 accumulators fill the 512 columns of tensor memory, which
// halves the operand bytes per MAC against a 128 x 128 tile (the kernel is L2-bandwidth bound
// otherwise).  Operand tiles go global -> shared with 16-byte cp.async copies placed directly in
// the UMMA no-swizzle K-major layout, three stages deep; k ranges are whole 128-byte chunks (the
// caller pads every window, padding bytes are 0 and add nothing).
namespace gram {

constexpr uint32_t CT = 256;                                // CTA tile (rows and columns of C)
constexpr uint32_t SUB_BYTES = 128 * BK;                    // one 128-row operand sub-tile: 16 KB
constexpr uint32_t STAGE_BYTES = 4 * SUB_BYTES;             // A0 A1 B0 B1
constexpr uint32_t NSTAGE = 3;
constexpr uint32_t SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024;
constexpr uint32_t TMEM_ALL = 512;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) k_gram_umma(const int8_t *__restrict__ X, size_t ld, uint32_t n,
    uint32_t k_lo, uint32_t k_hi, const ushort2 *__restrict__ tile_list, int32_t *__restrict__ C) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t s_bar[NSTAGE];
    __shared__ uint32_t s_tmem;
    // tiles of the upper block triangle, listed super-block by super-block: the CTAs resident at
    // one time share a few thousand rows of X, which then stay in L2
    const uint32_t bi = tile_list[blockIdx.x].x, bj = tile_list[blockIdx.x].y;
    const bool diag = bi == bj;
    const uint32_t i0 = bi * CT, j0 = bj * CT;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tiles = smem_u32(smem_raw) + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);

    if (tid == 0) {
        for (uint32_t st = 0; st < NSTAGE; st++) mbar_init(&s_bar[st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&s_tmem)), "r"(TMEM_ALL) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = s_tmem;

    const uint32_t r8 = lane & 7, kq = lane >> 3;
    const uint32_t nch = (k_hi - k_lo) / BK;  // whole chunks by contract
    // copies of one k chunk into a stage: sub-tiles A0 A1 (rows i0 ...) and, off the diagonal, B0 B1
    auto issue_copy = [&](uint32_t ch) {
        const uint32_t stage = tiles + (ch % NSTAGE) * STAGE_BYTES;
        const uint32_t k0 = k_lo + ch * BK;
        const uint32_t nsub = diag ? 2 : 4;
        for (uint32_t sub = 0; sub < nsub; sub++) {
            const uint32_t row0 = (sub < 2 ? i0 : j0) + (sub & 1) * 128;
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const uint32_t u = warp + 8 * p, g = u >> 1, kc = (u & 1) * 4 + kq;
                const uint32_t row = row0 + g * 8 + r8;
                const bool ok = row < n;  // rows past the end: zero fill (src-size 0)
                const int8_t *src = X + (size_t) (ok ? row : 0) * ld + k0 + kc * 16;
                cp_async16(stage + sub * SUB_BYTES + g * SBO + kc * LBO + r8 * 16, src, ok ? 16u : 0u);
            }
        }
    };
    for (uint32_t ch = 0; ch + 1 < NSTAGE; ch++) {
        if (ch < nch) issue_copy(ch);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (uint32_t ch = 0; ch < nch; ch++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(NSTAGE - 2) : "memory");  // chunk ch has landed
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t stage = tiles + (ch % NSTAGE) * STAGE_BYTES;
#pragma unroll
            for (uint32_t ks = 0; ks < BK / 32; ks++) {
#pragma unroll
                for (uint32_t a = 0; a < 2; a++) {
#pragma unroll
                    for (uint32_t b = 0; b < 2; b++) {
                        if (diag && a > b) continue;  // lower triangle of a diagonal block: never read
                        const uint32_t aaddr = stage + a * SUB_BYTES + ks * 2 * LBO;
                        const uint32_t baddr = stage + (diag ? b : 2 + b) * SUB_BYTES + ks * 2 * LBO;
                        umma_i8(tmem_acc + (a * 2 + b) * 128, umma_desc(aaddr), umma_desc(baddr),
                            (ch > 0 || ks > 0) ? 1u : 0u);
                    }
                }
            }
            umma_commit(&s_bar[ch % NSTAGE]);
        }
        // refill the stage read by the MMAs of chunk ch - 1 once they are complete
        const uint32_t nx = ch + NSTAGE - 1;
        if (nx < nch) {
            if (ch >= 1) mbar_wait(&s_bar[(ch - 1) % NSTAGE], ((ch - 1) / NSTAGE) & 1);
            issue_copy(nx);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (nch > 0) mbar_wait(&s_bar[(nch - 1) % NSTAGE], ((nch - 1) / NSTAGE) & 1);  // commits complete in order
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w owns TMEM lanes 32 (w % 4) .. + 31 and, of every accumulator, columns 64 (w / 4) ...
    for (uint32_t a = 0; a < 2; a++) {
        for (uint32_t b = 0; b < 2; b++) {
            if (diag && a > b) continue;
            const uint32_t row = i0 + a * 128 + 32 * (warp & 3) + lane;
#pragma unroll
            for (uint32_t half = 0; half < 2; half++) {
                const uint32_t col0 = 64 * (warp >> 2) + 32 * half;
                uint32_t v[32];
                const uint32_t taddr = tmem_acc + ((32 * (warp & 3)) << 16) + (a * 2 + b) * 128 + col0;
                if (nch > 0) {
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr) : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                } else {
#pragma unroll
                    for (int q = 0; q < 32; q++) v[q] = 0;
                }
                if (row < n) {
#pragma unroll
                    for (int q = 0; q < 32; q++) {
                        const uint32_t col = j0 + b * 128 + col0 + q;
                        if (col < n) C[(size_t) row * n + col] = (int32_t) v[q];
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(TMEM_ALL) : "memory");
    }
}

}  // namespace gram

// ---- G G^T through TMA: operand delivery by the bulk-tensor copy engine, warp-specialised
// One CTA per 256 x 256 block of C as above (four 128 x 128 accumulators = all 512 TMEM columns).
// Warp 0 (one lane) is the PRODUCER: per 128-byte K chunk it arms the stage's "full" barrier with the
// byte count and issues four cp.async.bulk.tensor.2d loads (A0 A1 B0 B1, 128 rows x 128 bytes each,
// SWIZZLE_128B: the canonical K-major UMMA layout, rows past n filled with zeros by the copy engine).
// Warp 1 (one lane) is the MMA ISSUER: waits for "full", issues 16 tcgen05.mma (kind::i8, M 128, N 128,
// K 32) and tcgen05.commit's to the stage's "empty" barrier, which the producer waits on before it
// refills.  No __syncthreads and no thread-issued copies in the main loop.  Diagonal blocks load and
// multiply like the others (their lower half is simply not stored), so that all resident CTAs walk K
// in step and a chunk fetched by one is still in L2 for the others.
namespace gram_tma {

constexpr uint32_t CT = 256, SUB_BYTES = 128 * 128, STAGE_BYTES = 4 * SUB_BYTES, NSTAGE = 3;
constexpr uint32_t SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024;
constexpr uint32_t TMEM_ALL = 512;
constexpr uint32_t SPIN_MAX = 1u << 28;  // a lost copy must not hang the GPU: trap instead

__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > SPIN_MAX) asm volatile("trap;");
    }
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint64_t *bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor): 8-row x 128-byte atoms, 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t) ((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t) 1 << 16;                      // leading byte offset: unused for swizzled K-major
    d |= (uint64_t) ((1024u >> 4) & 0x3fffu) << 32;  // stride byte offset between 8-row groups
    d |= (uint64_t) 1 << 46;                      // descriptor version
    d |= (uint64_t) 2 << 61;                      // SWIZZLE_128B
    return d;
}

__global__ void __launch_bounds__(THREADS, 1) k_gram_tma(const __grid_constant__ CUtensorMap tmap, uint32_t n,
    uint32_t k_lo, uint32_t k_hi, const ushort2 *__restrict__ tile_list, int32_t *__restrict__ C) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t s_full[NSTAGE], s_empty[NSTAGE], s_done;
    __shared__ uint32_t s_tmem;
    const uint32_t bi = tile_list[blockIdx.x].x, bj = tile_list[blockIdx.x].y;
    const bool diag = bi == bj;
    const uint32_t i0 = bi * CT, j0 = bj * CT;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tiles = smem_u32(smem_raw) + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    if (tid == 0) {
        for (uint32_t st = 0; st < NSTAGE; st++) {
            mbar_init(&s_full[st], 1);
            mbar_init(&s_empty[st], 1);
        }
        mbar_init(&s_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&s_tmem)), "r"(TMEM_ALL) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = s_tmem;
    const uint32_t nch = (k_hi - k_lo) / 128;  // whole 128-byte chunks by contract

    if (warp == 0 && lane == 0) {
        // ---- producer
        for (uint32_t ch = 0; ch < nch; ch++) {
            const uint32_t st = ch % NSTAGE, round = ch / NSTAGE;
            if (round > 0) mbar_wait_bounded(&s_empty[st], (round - 1) & 1);
            mbar_expect_tx(&s_full[st], STAGE_BYTES);
            const uint32_t stage = tiles + st * STAGE_BYTES;
            const int32_t k0 = (int32_t) (k_lo + ch * 128);
            tma_load_2d(stage + 0 * SUB_BYTES, &tmap, &s_full[st], k0, (int32_t) i0);
            tma_load_2d(stage + 1 * SUB_BYTES, &tmap, &s_full[st], k0, (int32_t) (i0 + 128));
            tma_load_2d(stage + 2 * SUB_BYTES, &tmap, &s_full[st], k0, (int32_t) j0);
            tma_load_2d(stage + 3 * SUB_BYTES, &tmap, &s_full[st], k0, (int32_t) (j0 + 128));
        }
    } else if (warp == 1 && lane == 0) {
        // ---- MMA issuer
        for (uint32_t ch = 0; ch < nch; ch++) {
            const uint32_t st = ch % NSTAGE, round = ch / NSTAGE;
            mbar_wait_bounded(&s_full[st], round & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t stage = tiles + st * STAGE_BYTES;
#pragma unroll
            for (uint32_t ks = 0; ks < 4; ks++) {  // 32 bytes of K per instruction, inside the 128-byte swizzle row
#pragma unroll
                for (uint32_t a = 0; a < 2; a++) {
#pragma unroll
                    for (uint32_t b = 0; b < 2; b++) {
                        umma_i8(tmem_acc + (a * 2 + b) * 128, umma_desc_sw128(stage + a * SUB_BYTES + ks * 32),
                            umma_desc_sw128(stage + (2 + b) * SUB_BYTES + ks * 32), (ch > 0 || ks > 0) ? 1u : 0u);
                    }
                }
            }
            umma_commit(&s_empty[st]);  // arrives when these MMAs have read the stage
        }
        umma_commit(&s_done);           // ... and when every MMA of the block is complete
        if (nch > 0) mbar_wait_bounded(&s_done, 0);
    }
    // ---- epilogue: all warps; warp w owns TMEM lanes 32 (w % 4) .. + 31 and columns 64 (w / 4) ... of every accumulator.
    // The warps without a role sleep in the block barrier during the main loop (no polling); the MMA
    // issuer arrives once the last commit has completed.
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (uint32_t a = 0; a < 2; a++) {
        for (uint32_t b = 0; b < 2; b++) {
            if (diag && a > b) continue;  // lower half of a diagonal block: never read
            const uint32_t row = i0 + a * 128 + 32 * (warp & 3) + lane;
#pragma unroll
            for (uint32_t half = 0; half < 2; half++) {
                const uint32_t col0 = 64 * (warp >> 2) + 32 * half;
                uint32_t v[32];
                const uint32_t taddr = tmem_acc + ((32 * (warp & 3)) << 16) + (a * 2 + b) * 128 + col0;
                if (nch > 0) {
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr) : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                } else {
#pragma unroll
                    for (int q = 0; q < 32; q++) v[q] = 0;
                }
                if (row < n) {
#pragma unroll
                    for (int q = 0; q < 32; q++) {
                        const uint32_t col = j0 + b * 128 + col0 + q;
                        if (col < n) C[(size_t) row * n + col] = (int32_t) v[q];
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(TMEM_ALL) : "memory");
    }
}

// tensor map of X [n rows][ld bytes] for 128-row x 128-byte boxes, SWIZZLE_128B; the driver entry point
// is looked up at run time (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline bool make_tensor_map(CUtensorMap *map, const int8_t *X, size_t ld, uint32_t n) {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess
            || q != cudaDriverEntryPointSuccess || p == nullptr) {
            cudaGetLastError();
            return false;
        }
        fn = (EncodeTiledFn) p;
    }
    const cuuint64_t dims[2] = { (cuuint64_t) ld, (cuuint64_t) n };
    const cuuint64_t strides[1] = { (cuuint64_t) ld };
    const cuuint32_t box[2] = { 128, 128 }, elem[2] = { 1, 1 };
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *) X, dims, strides, box, elem,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace gram_tma

}  // namespace tc

// number of sites at which samples lo < hi differ, from the contraction's output
__device__ __forceinline__ int64_t pair_different(const int32_t *same, uint32_t n, uint32_t lo, uint32_t hi,
    uint32_t ncols, int biallelic) {
    const int64_t c = same[(size_t) lo * n + hi];
    if (biallelic) return (int64_t) same[(size_t) lo * n + lo] + (int64_t) same[(size_t) hi * n + hi] - 2 * c;
    return (int64_t) ncols - c;
}

// D[a][b] = sum over j in set a, k in set b of (sites - same[j][k]) (j != k), count- and
// span-normalised (trees.c:8876-8899, 1920-1934).  `same` holds blocks with row block <= col
// block only: read the transposed entry otherwise.
__global__ void k_divmat_finish(const int32_t *same, uint32_t n, const uint32_t *set_off,
    uint32_t nsets, const double *set_size, uint32_t ncols, int biallelic, double span,
    int span_normalise, double *D) {
    typedef cub::BlockReduce<double, TB> BR;
    __shared__ typename BR::TempStorage tmp;
    const uint32_t a = blockIdx.y, b = blockIdx.x;
    if (a > b) return;
    const uint32_t r0 = set_off[a], r1 = set_off[a + 1], c0 = set_off[b], c1 = set_off[b + 1];
    const uint32_t nr = r1 - r0, nc = c1 - c0;
    double sum = 0.0;
    for (size_t t = threadIdx.x; t < (size_t) nr * nc; t += TB) {
        uint32_t j = r0 + (uint32_t) (t / nc), k = c0 + (uint32_t) (t % nc);
        if (j == k) continue;
        uint32_t lo = j < k ? j : k, hi = j < k ? k : j;
        // entries of the upper block triangle: (lo, hi) is stored iff block(lo) <= block(hi): always
        sum += (double) pair_different(same, n, lo, hi, ncols, biallelic);
    }
    double tot = BR(tmp).Sum(sum);
    if (threadIdx.x == 0) {
        double denom = a == b ? set_size[a] * (set_size[a] - 1) : set_size[a] * set_size[b];
        if (!(a == b && denom == 0)) tot /= denom;
        if (span_normalise) tot /= span;
        D[(size_t) a * nsets + b] = tot;
        D[(size_t) b * nsets + a] = tot;
    }
}

// every set a single sample (the default of divergence_matrix)
__global__ void k_divmat_pairs(const int32_t *same, uint32_t n, uint32_t ncols, int biallelic, double span,
    int span_normalise, double *D) {
    const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t) n * n) return;
    const uint32_t j = (uint32_t) (t / n), k = (uint32_t) (t % n);
    double v = 0.0;
    if (j != k) {
        const uint32_t lo = j < k ? j : k, hi = j < k ? k : j;
        v = (double) pair_different(same, n, lo, hi, ncols, biallelic);
        if (span_normalise) v /= span;
    }
    D[t] = v;
}

struct Temp {
    void *p = nullptr;
    size_t cap = 0;
    ~Temp() { if (p) cudaFree(p); }
    void *need(size_t bytes) {
        if (bytes > cap) {
            if (p) cudaFree(p);
            cap = bytes + 1024;
            TSKB_CK(cudaMalloc(&p, cap));
        }
        return p;
    }
};

// genotypes of sites [s0, s1) for the listed samples into G (caller-zeroed), any strides
void decode_sites(const Plan &P, const int32_t *d_samples, uint32_t n, uint32_t s0, uint32_t s1,
    uint32_t options, int8_t *G, size_t stride_site, size_t stride_sample,
    const uint32_t *site_col = nullptr) {
    if (s1 <= s0 || n == 0) return;
    cudaStream_t s = P.stream;
    const uint32_t N = (uint32_t) P.N;
    DevArray<int32_t> col;
    col.alloc(N);
    k_fill_i32<<<grid_for(N, TB), TB, 0, s>>>(col.p, N, -1);
    k_set_cols<<<grid_for(n, TB), TB, 0, s>>>(d_samples, n, col.p);
    TSKB_CK_LAUNCH();
    if (!(options & TSKB_ISOLATED_NOT_MISSING)) {
        k_mark_isolated<<<grid_for(n, 128), 128, 0, s>>>(d_samples, n, P.coff.p, P.csr_left.p,
            P.csr_right.p, P.pm_off.p, P.pm_left.p, P.pm_right.p, P.site_pos.p, s0, s1, P.L, G, stride_site,
            stride_sample);
        TSKB_CK_LAUNCH();
    }
    // frontier buffers; a batch of sites is re-run with half the sites if a layer overflows
    const uint32_t cap = 1u << 26;
    DevArray<uint2> fa, fb;
    DevArray<uint32_t> cnt;
    DevArray<int> ovf;
    fa.alloc(cap); fb.alloc(cap); cnt.alloc(2); ovf.alloc(1);
    uint32_t batch = std::max<uint32_t>(1, std::min<uint32_t>(s1 - s0, cap / std::max<uint32_t>(1, n / 4)));
    uint32_t b0 = s0;
    while (b0 < s1) {
        const uint32_t b1 = std::min(s1, b0 + batch);
        bool overflow = false;
        for (uint32_t r = 0; r < P.max_muts_per_site && !overflow; r++) {
            uint32_t h_cnt = 0;
            TSKB_CK(cudaMemsetAsync(cnt.p, 0, 2 * sizeof(uint32_t), s));
            TSKB_CK(cudaMemsetAsync(ovf.p, 0, sizeof(int), s));
            k_frontier_init<<<grid_for(b1 - b0, TB), TB, 0, s>>>(r, b0, b1, P.site_moff.p, P.mut_node.p,
                fa.p, cnt.p);
            TSKB_CK_LAUNCH();
            TSKB_CK(cudaMemcpyAsync(&h_cnt, cnt.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            TSKB_CK(cudaStreamSynchronize(s));
            uint2 *cur = fa.p, *nxt = fb.p;
            uint32_t guard = 0;
            while (h_cnt > 0) {
                TSKB_CK(cudaMemsetAsync(cnt.p + 1, 0, sizeof(uint32_t), s));
                k_expand<<<grid_for(h_cnt, TB), TB, 0, s>>>(cur, h_cnt, r, s0, P.site_pos.p,
                    P.site_moff.p, P.mut_allele.p, col.p, P.pm_off.p, P.pm_left.p, P.pm_right.p, P.pm_pmax.p,
                    P.pm_child.p, G, stride_site, stride_sample, site_col, nxt, cap, cnt.p + 1, ovf.p);
                TSKB_CK_LAUNCH();
                int h_ovf = 0;
                TSKB_CK(cudaMemcpyAsync(&h_cnt, cnt.p + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
                TSKB_CK(cudaMemcpyAsync(&h_ovf, ovf.p, sizeof(int), cudaMemcpyDeviceToHost, s));
                TSKB_CK(cudaStreamSynchronize(s));
                if (h_ovf) {
                    overflow = true;
                    break;
                }
                std::swap(cur, nxt);
                if (++guard > (1u << 20)) throw (int) TSKB_ERR_BAD_PARAM_VALUE;  // cyclic input
            }
        }
        if (overflow) {
            if (batch == 1) throw (int) TSKB_ERR_NO_MEMORY;
            batch = std::max<uint32_t>(1, batch / 2);
            continue;  // redo this batch: rounds rewrite their genotypes in order
        }
        b0 = b1;
    }
}

// tsk_treeseq_divergence_matrix argument checks (trees.c:8901-8990), same precedence
int check_divmat_args(const Plan &P, uint64_t &nsets, const uint64_t *&sizes, const int32_t *&sets,
    std::vector<uint64_t> &tmp_sizes, uint64_t &num_windows, const double *&windows,
    double *default_windows, uint32_t &options, uint64_t &total) {
    bool site = options & TSKB_STAT_SITE, branch = options & TSKB_STAT_BRANCH;
    if (options & TSKB_STAT_NODE) return TSKB_ERR_UNSUPPORTED_STAT_MODE;
    if (!(site || branch)) {
        site = true;
        options |= TSKB_STAT_SITE;
    }
    if (site + branch > 1) return TSKB_ERR_MULTIPLE_STAT_MODES;
    if (options & TSKB_STAT_POLARISED) return TSKB_ERR_STAT_POLARISED_UNSUPPORTED;
    if (windows == nullptr) {
        num_windows = 1;
        default_windows[0] = 0;
        default_windows[1] = P.L;
        windows = default_windows;
    } else {
        if (num_windows < 1) return TSKB_ERR_BAD_NUM_WINDOWS;
        if (windows[0] < 0 || windows[num_windows] > P.L) return TSKB_ERR_BAD_WINDOWS;
        for (uint64_t j = 0; j < num_windows; j++) {
            if (windows[j] >= windows[j + 1]) return TSKB_ERR_BAD_WINDOWS;
        }
    }
    if (sets == nullptr) {
        sets = P.samples.data();
        if (sizes == nullptr) nsets = P.num_samples;
    }
    if (sizes == nullptr) {
        tmp_sizes.assign(nsets, 1);
        sizes = tmp_sizes.data();
    }
    std::vector<int32_t> seen(P.N, -1);
    total = 0;
    uint64_t i = 0;
    for (uint64_t j = 0; j < nsets; j++) {
        total += sizes[j];
        for (uint64_t k = 0; k < sizes[j]; k++, i++) {
            int32_t u = sets[i];
            if (u < 0 || u >= (int32_t) P.N) return TSKB_ERR_NODE_OUT_OF_BOUNDS;
            if (P.sample_index_map[u] == -1) return TSKB_ERR_BAD_SAMPLES;
            if (seen[u] != -1) return TSKB_ERR_DUPLICATE_SAMPLE;
            seen[u] = (int32_t) j;
        }
    }
    return 0;
}

// Divergence matrix, branch mode (tsk_treeseq_divergence_matrix_branch, trees.c:8579-8676).
// The reference adds, per tree and per pair of samples (u, v), the path length between them
// ((t_mrca - t_u) + (t_mrca - t_v), or the two distances to the roots when they are in different
// subtrees) times the tree's span, then divides by the number of pairs (trees.c:8876-8899).
// The path between u and v consists of exactly the branches that have one of the two below them,
// so entry (j, k) is the branch-mode `divergence` statistic of sample sets j and k
// (trees.c:4221-4264 with the branch summary of trees.c:1944-1972): sum over branches of
// length * [x_j (n_k - x_k) + (n_j - x_j) x_k] / (n_j n_k), and 2 x_j (n_j - x_j) / (n_j (n_j - 1))
// on the diagonal.  The matrix is therefore computed by the sweep engine itself, in blocks of
// BLK x BLK sets (2 * BLK state columns per sweep), over the windows padded to [0, L].
constexpr uint32_t BLK = 4;

int run_branch_divergence_matrix(const Plan &P, uint64_t nsets, const uint64_t *sizes, const int32_t *sets,
    uint64_t num_windows, const double *windows, uint32_t options, double *result) {
    const uint32_t ns = (uint32_t) nsets, W = (uint32_t) num_windows;
    memset(result, 0, (size_t) W * ns * ns * sizeof(double));
    if (ns == 0) return 0;
    // the sweep engine takes windows covering the whole genome (trees.c:2070); divergence_matrix
    // windows need not (trees.c:8943): pad, and drop the padding windows from the output
    std::vector<double> win;
    const uint32_t lead = windows[0] > 0 ? 1 : 0;
    if (lead) win.push_back(0.0);
    win.insert(win.end(), windows, windows + W + 1);
    if (windows[W] < P.L) win.push_back(P.L);
    const uint32_t Wp = (uint32_t) win.size() - 1;
    std::vector<uint64_t> off(ns + 1, 0);
    for (uint32_t a = 0; a < ns; a++) off[a + 1] = off[a] + sizes[a];
    const uint32_t nblk = (ns + BLK - 1) / BLK;
    std::vector<uint64_t> b_sizes;
    std::vector<int32_t> b_sets, tuples;
    std::vector<uint32_t> tup_a, tup_b;
    std::vector<double> out;
    for (uint32_t ba = 0; ba < nblk; ba++) {
        for (uint32_t bb = ba; bb < nblk; bb++) {
            b_sizes.clear(); b_sets.clear(); tuples.clear(); tup_a.clear(); tup_b.clear();
            const uint32_t a0 = ba * BLK, a1 = std::min(ns, a0 + BLK);
            const uint32_t c0 = bb * BLK, c1 = std::min(ns, c0 + BLK);
            auto push_set = [&](uint32_t a) {
                b_sizes.push_back(sizes[a]);
                b_sets.insert(b_sets.end(), sets + off[a], sets + off[a + 1]);
            };
            for (uint32_t a = a0; a < a1; a++) push_set(a);
            if (bb != ba) {
                for (uint32_t c = c0; c < c1; c++) push_set(c);
            }
            for (uint32_t a = a0; a < a1; a++) {
                for (uint32_t c = std::max(a, c0); c < c1; c++) {
                    // a singleton set has no pairs: the reference leaves 0 on the diagonal
                    // (trees.c:8888-8891) where the statistic would be 0/0
                    if (a == c && sizes[a] < 2) continue;
                    tuples.push_back((int32_t) (a - a0));
                    tuples.push_back((int32_t) (bb != ba ? (a1 - a0) + (c - c0) : c - a0));
                    tup_a.push_back(a);
                    tup_b.push_back(c);
                }
            }
            const uint32_t M = (uint32_t) tup_a.size();
            if (M == 0) continue;
            out.assign((size_t) Wp * M, 0.0);
            StatSpec sp = {};
            sp.stat_id = STAT_DIVERGENCE;
            sp.K = (uint32_t) b_sizes.size();
            sp.M = M;
            sp.tuple = 2;
            sp.sizes = b_sizes.data();
            sp.sets = b_sets.data();
            sp.sets_on_device = false;
            sp.indexes = tuples.data();
            sp.W = Wp;
            sp.windows = win.data();
            sp.options = TSKB_STAT_BRANCH | (options & TSKB_STAT_SPAN_NORMALISE);
            sp.result = out.data();
            sp.result_on_device = false;
            int ret = run_sample_count_stat(&P, sp);
            if (ret != 0) return ret;
            for (uint32_t w = 0; w < W; w++) {
                double *D = result + (size_t) w * ns * ns;
                for (uint32_t m = 0; m < M; m++) {
                    const double v = out[(size_t) (w + lead) * M + m];
                    D[(size_t) tup_a[m] * ns + tup_b[m]] = v;
                    D[(size_t) tup_b[m] * ns + tup_a[m]] = v;
                }
            }
        }
    }
    return 0;
}

}  // namespace

int run_genotype_matrix(const Plan *plan, const int32_t *samples, uint64_t num_samples,
    uint32_t options, int8_t *genotypes, uint64_t first_site, uint64_t num_sites) {
    std::lock_guard<std::mutex> lock(plan->mu);
    const Plan &P = *plan;
    TSKB_CK(cudaSetDevice(P.device));
    cudaStream_t s = P.stream;
    if (first_site > P.S || num_sites > P.S - first_site) return -205;  // TSK_ERR_SITE_OUT_OF_BOUNDS
    const uint32_t S0 = (uint32_t) first_site, S = (uint32_t) num_sites;
    DevArray<int32_t> d_s;
    const int32_t *ds = P.d_samples.p;
    uint32_t n = P.num_samples;
    if (samples != nullptr) {
        for (uint64_t j = 0; j < num_samples; j++) {
            if (samples[j] < 0 || samples[j] >= (int32_t) P.N) return TSKB_ERR_NODE_OUT_OF_BOUNDS;
        }
        d_s.upload(samples, num_samples, s);
        ds = d_s.p;
        n = (uint32_t) num_samples;
    }
    if (S == 0 || n == 0) return 0;
    DevArray<int8_t> G;
    G.alloc((size_t) S * n);
    TSKB_CK(cudaMemsetAsync(G.p, 0, (size_t) S * n, s));
    decode_sites(P, ds, n, S0, S0 + S, options, G.p, n, 1);
    staged_download(P.device, genotypes, G.p, (size_t) S * n, s);
    TSKB_CK(cudaStreamSynchronize(s));
    return 0;
}

int run_divergence_matrix(const Plan *plan, uint64_t nsets, const uint64_t *sizes, const int32_t *sets,
    uint64_t num_windows, const double *windows, uint32_t options, double *result) {
    const Plan &P = *plan;
    std::vector<uint64_t> tmp_sizes;
    double default_windows[2];
    uint64_t total = 0;
    int ret = check_divmat_args(P, nsets, sizes, sets, tmp_sizes, num_windows, windows, default_windows,
        options, total);
    if (ret != 0) return ret;
    if (options & TSKB_STAT_BRANCH) {
        if (P.time_uncalibrated && !(options & TSKB_STAT_ALLOW_TIME_UNCALIBRATED)) {
            return TSKB_ERR_TIME_UNCALIBRATED;
        }
        return run_branch_divergence_matrix(P, nsets, sizes, sets, num_windows, windows, options, result);
    }
    std::lock_guard<std::mutex> lock(P.mu);
    TSKB_CK(cudaSetDevice(P.device));
    cudaStream_t s = P.stream;
    const uint32_t n = (uint32_t) total, W = (uint32_t) num_windows, ns = (uint32_t) nsets;
    if (n == 0) {  // (every window's matrix is otherwise written whole by the read-back below)
        memset(result, 0, (size_t) W * ns * ns * sizeof(double));
        return 0;
    }
    // sites covered by the windows
    // (a plan staged for a genome range contracts the sites inside its range only: the per-range
    // partial matrices of a sharded call add up to the whole, exactly -- they are integer counts)
    auto site_index = [&](double x) {
        const uint32_t i = (uint32_t) (std::lower_bound(P.h_site_pos.begin(), P.h_site_pos.end(), x) - P.h_site_pos.begin());
        return std::min(std::max(i, P.site_lo), P.site_hi);
    };
    const uint32_t S0 = site_index(windows[0]), S1 = site_index(windows[W]);
    // genotype columns: every window starts on a 128-byte boundary and is padded to whole 128-byte
    // chunks with zeros (ancestral everywhere: the same for every pair), so k ranges never need masks
    std::vector<uint32_t> w_lo(W + 1), col_lo(W + 1);
    for (uint32_t w = 0; w <= W; w++) w_lo[w] = site_index(windows[w]);
    uint64_t cols = 0;
    for (uint32_t w = 0; w < W; w++) {
        col_lo[w] = (uint32_t) cols;
        cols += ((uint64_t) (w_lo[w + 1] - w_lo[w]) + 127) / 128 * 128;
    }
    col_lo[W] = (uint32_t) cols;
    if (cols >= 0xffffff00ull) return TSKB_ERR_UNSUPPORTED;
    const size_t ld = (size_t) cols + 128;
    std::vector<uint32_t> h_site_col(S1 - S0);
    for (uint32_t w = 0; w < W; w++) {
        for (uint32_t sidx = w_lo[w]; sidx < w_lo[w + 1]; sidx++) h_site_col[sidx - S0] = col_lo[w] + (sidx - w_lo[w]);
    }
    DevArray<uint32_t> d_site_col;
    d_site_col.upload(h_site_col.data(), h_site_col.size(), s);
    DevArray<int32_t> d_sets;
    d_sets.upload(sets, n, s);
    DevArray<int8_t> X;
    X.alloc((size_t) n * ld);
    TSKB_CK(cudaMemsetAsync(X.p, 0, (size_t) n * ld, s));
    // the reference decodes with TSK_ISOLATED_NOT_MISSING here (trees.c:8775): genotypes are >= 0
    TSKB_CK(cudaEventRecord(P.ev[0], s));
    decode_sites(P, d_sets.p, n, S0, S1, TSKB_ISOLATED_NOT_MISSING, X.p, 1, ld, d_site_col.p);
    TSKB_CK(cudaEventRecord(P.ev[1], s));
    float gemm_ms = 0, finish_ms = 0;
    std::vector<uint32_t> h_off(ns + 1, 0);
    std::vector<double> h_size(ns);
    for (uint32_t a = 0; a < ns; a++) {
        h_off[a + 1] = h_off[a] + (uint32_t) sizes[a];
        h_size[a] = (double) sizes[a];
    }
    DevArray<uint32_t> d_off;
    DevArray<double> d_size, d_D;
    DevArray<int32_t> same;
    d_off.upload(h_off.data(), ns + 1, s);
    d_size.upload(h_size.data(), ns, s);
    d_D.alloc((size_t) ns * ns);
    same.alloc((size_t) n * n);
    // TSKB_MATRIX=legacy: mma.sync one-hot path; =onehot: tcgen05 one-hot path even for biallelic
    // data (both kept for A/B measurements and as cross-checks in the tests)
    const char *mode_env = getenv("TSKB_MATRIX");
    const bool use_legacy = mode_env != nullptr && mode_env[0] == 'l';
    const bool biallelic = P.max_alleles_per_site <= 2 && !use_legacy
                           && !(mode_env != nullptr && mode_env[0] == 'o');
    // operand delivery: TMA (default) or thread-issued cp.async (TSKB_MATRIX=cpasync, the earlier kernel)
    CUtensorMap tmap;
    bool use_tma = biallelic && !(mode_env != nullptr && mode_env[0] == 'c');
    if (use_tma) use_tma = tc::gram_tma::make_tensor_map(&tmap, X.p, ld, n);
    if (biallelic && use_tma) {
        TSKB_CK(cudaFuncSetAttribute(tc::gram_tma::k_gram_tma, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int) tc::gram_tma::SMEM_BYTES));
    } else if (biallelic) {
        TSKB_CK(cudaFuncSetAttribute(tc::gram::k_gram_umma, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int) tc::gram::SMEM_BYTES));
    } else if (!use_legacy) {
        TSKB_CK(cudaFuncSetAttribute(tc::k_same_umma, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int) tc::SMEM_BYTES));
    }
    std::vector<ushort2> h_tiles;
    DevArray<ushort2> d_tiles;
    if (biallelic) {
        const uint32_t nb = (n + tc::gram::CT - 1) / tc::gram::CT, SB = 12;  // 12 x 12 tiles ~ one wave
        for (uint32_t si = 0; si < nb; si += SB) {
            for (uint32_t sj = si; sj < nb; sj += SB) {
                for (uint32_t bi = si; bi < std::min(nb, si + SB); bi++) {
                    for (uint32_t bj = std::max(bi, sj); bj < std::min(nb, sj + SB); bj++) {
                        h_tiles.push_back(make_ushort2((unsigned short) bi, (unsigned short) bj));
                    }
                }
            }
        }
        d_tiles.upload(h_tiles.data(), h_tiles.size(), s);
    }
    for (uint32_t w = 0; w < W; w++) {
        const uint32_t k_lo = col_lo[w], k_hi = col_lo[w + 1];  // padded: whole 128-byte chunks
        TSKB_CK(cudaMemsetAsync(same.p, 0, (size_t) n * n * sizeof(int32_t), s));
        if (k_hi > k_lo) {
            if (biallelic && use_tma) {
                tc::gram_tma::k_gram_tma<<<(unsigned) h_tiles.size(), tc::THREADS, tc::gram_tma::SMEM_BYTES, s>>>(
                    tmap, n, k_lo, k_hi, d_tiles.p, same.p);
                TSKB_CK_LAUNCH();
            } else if (biallelic) {
                tc::gram::k_gram_umma<<<(unsigned) h_tiles.size(), tc::THREADS, tc::gram::SMEM_BYTES, s>>>(
                    X.p, ld, n, k_lo, k_hi, d_tiles.p, same.p);
                TSKB_CK_LAUNCH();
            } else if (use_legacy) {
                const uint32_t nb = (n + GM - 1) / GM;
                for (uint32_t a = 0; a < P.max_alleles_per_site; a++) {
                    k_same_gemm<<<dim3(nb, nb), TB, 0, s>>>(X.p, ld, n, k_lo, k_hi, (int) a, same.p);
                    TSKB_CK_LAUNCH();
                }
            } else {
                const uint32_t nb = (n + tc::TM - 1) / tc::TM;
                tc::k_same_umma<<<dim3(nb, nb), tc::THREADS, tc::SMEM_BYTES, s>>>(X.p, ld, n, k_lo, k_hi,
                    P.max_alleles_per_site, same.p);
                TSKB_CK_LAUNCH();
            }
        }
        TSKB_CK(cudaEventRecord(P.ev[2], s));
        // one-hot paths: different = columns - same (padding columns count as same for every pair);
        // biallelic path: different = C[i][i] + C[j][j] - 2 C[i][j]
        const uint32_t ncols = k_hi - k_lo;
        const int span = (options & TSKB_STAT_SPAN_NORMALISE) ? 1 : 0;
        if (ns == n) {
            k_divmat_pairs<<<grid_for((size_t) n * n, TB), TB, 0, s>>>(same.p, n, ncols, biallelic ? 1 : 0,
                windows[w + 1] - windows[w], span, d_D.p);
        } else {
            k_divmat_finish<<<dim3(ns, ns), TB, 0, s>>>(same.p, n, d_off.p, ns, d_size.p, ncols,
                biallelic ? 1 : 0, windows[w + 1] - windows[w], span, d_D.p);
        }
        TSKB_CK_LAUNCH();
        staged_download(P.device, result + (size_t) w * ns * ns, d_D.p, (size_t) ns * ns * sizeof(double), s);
        TSKB_CK(cudaEventRecord(P.ev[3], s));
        TSKB_CK(cudaStreamSynchronize(s));
        float ms = 0;
        TSKB_CK(cudaEventElapsedTime(&ms, w == 0 ? P.ev[1] : P.ev[4], P.ev[2]));
        gemm_ms += ms;
        TSKB_CK(cudaEventElapsedTime(&ms, P.ev[2], P.ev[3]));
        finish_ms += ms;
        TSKB_CK(cudaEventRecord(P.ev[4], s));
    }
    TSKB_CK(cudaStreamSynchronize(s));
    // phase times of the last matrix call: 0 decode, 1 contraction, 2 normalise + copy to host;
    // 7 = alleles looped per site
    float decode_ms = 0;
    TSKB_CK(cudaEventElapsedTime(&decode_ms, P.ev[0], P.ev[1]));
    for (double &v : P.stats.last_kernel_ms) v = 0;
    P.stats.last_kernel_ms[0] = decode_ms;
    P.stats.last_kernel_ms[1] = gemm_ms;
    P.stats.last_kernel_ms[2] = finish_ms;
    P.stats.last_kernel_ms[7] = biallelic ? 1 : P.max_alleles_per_site;
    P.stats.last_call_ms = decode_ms + gemm_ms + finish_ms;
    return 0;
}

}  // namespace tskb
