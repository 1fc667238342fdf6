// exchange.cu -- summing the ranks' per-window partials over NVLink peer memory, without NCCL.
//
// The only exchange on the genome-sharded path is the sum of W x M doubles per statistic (a few KB to
// a few MB).  Through NCCL that costs a host-side launch per call plus the collective's own latency,
// more than the transfer itself.  Here every rank owns ONE receive slab (cudaMalloc, exported once as
// a CUDA IPC handle; the peers map it with cudaIpcOpenMemHandle from their own device, which also
// enables peer access), laid out as
//
//     [parity 0 | parity 1] x [source rank] x [capacity] doubles, then [parity] x [source rank] flags.
//
// A call PUSHES this rank's partial into slot [rank] of every peer's slab (plain stores over NVLink),
// publishes the call's epoch number in the peer's flag after a system-scope fence, waits (device side,
// bounded by a wall-clock limit) until its own flags have all reached the epoch, and adds the world's
// slots in rank order -- the same order on every rank, so the sum is bit-identical everywhere -- and
// span-normalises (trees.c:1920-1934 divides after accumulation).  Small results (the usual case) do
// all four steps in ONE single-CTA kernel; large ones in a grid-wide push, a signal/wait kernel and a
// grid-wide sum.  Slabs are double-buffered by epoch parity: a rank can be at most one call ahead of
// a peer, because finishing a call needs every peer's partial of that call.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstring>
#include <new>

#include "plan.cuh"

struct tskb_exchange {
    int device = 0;
    uint32_t world = 1, rank = 0;
    uint64_t capacity = 0;              // doubles per (parity, source rank) slot
    char *slab = nullptr;               // this rank's receive slab (cudaMalloc)
    std::vector<char *> peer_slab;      // [world]: mapped bases (own slab at [rank])
    std::vector<bool> mapped;           // opened by cudaIpcOpenMemHandle (to close)
    double **d_peer_recv = nullptr;     // device table [2][world]
    uint32_t **d_peer_flags = nullptr;  // device table [2][world]
    int *h_timed_out = nullptr;         // mapped pinned host word written by a wait that gave up
    uint32_t epoch = 0;
    cudaStream_t stream = nullptr;      // used when a call names no engine
    bool connected = false;
    std::mutex mu;
};

namespace tskb {
namespace {

constexpr int TBX = 256;
constexpr int FUSED_THREADS = 1024;
constexpr uint64_t FUSED_MAX = 1u << 15;              // doubles: above this the grid-wide kernels
constexpr uint64_t WAIT_NS = 20ull * 1000000000ull;    // a peer that died must not hang this GPU

__device__ __forceinline__ uint64_t now_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void publish(uint32_t *flag, uint32_t epoch) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
}

__device__ __forceinline__ void await(const uint32_t *flag, uint32_t epoch, int *timed_out) {
    uint32_t v = 0;
    const uint64_t t0 = now_ns();
    while (true) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int32_t) (v - epoch) >= 0) break;
        if (now_ns() - t0 > WAIT_NS) {
            *(volatile int *) timed_out = 1;
            break;
        }
        __nanosleep(200);
    }
}

__device__ __forceinline__ double span_of(const double *spans, uint64_t i, uint64_t stride, uint64_t n) {
    return spans[(i / stride) % n];
}

// everything in one CTA: push, publish, wait, rank-ordered sum
__global__ void __launch_bounds__(FUSED_THREADS) k_exchange_fused(const double *__restrict__ local, uint64_t count,
    uint64_t capacity, uint32_t world, uint32_t rank, double *const *__restrict__ peer_recv,
    uint32_t *const *__restrict__ peer_flags, uint32_t epoch, const double *__restrict__ spans, uint64_t span_stride,
    uint64_t span_count, double *out, int *timed_out) {
    for (uint64_t i = threadIdx.x; i < count; i += blockDim.x) {
        const double v = local[i];
        for (uint32_t p = 0; p < world; p++) peer_recv[p][(uint64_t) rank * capacity + i] = v;
    }
    __syncthreads();
    if (threadIdx.x < world) {
        publish(peer_flags[threadIdx.x] + rank, epoch);
        await(peer_flags[rank] + threadIdx.x, epoch, timed_out);
    }
    __syncthreads();
    const double *recv = peer_recv[rank];
    for (uint64_t i = threadIdx.x; i < count; i += blockDim.x) {
        double s = 0.0;
        for (uint32_t r = 0; r < world; r++) s += __ldcv(recv + (uint64_t) r * capacity + i);  // written by peers
        if (spans != nullptr) s /= span_of(spans, i, span_stride, span_count);
        out[i] = s;
    }
}

__global__ void k_exchange_push(const double *__restrict__ local, uint64_t count, uint64_t capacity, uint32_t world,
    uint32_t rank, double *const *__restrict__ peer_recv) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double v = local[i];
    for (uint32_t p = 0; p < world; p++) peer_recv[p][(uint64_t) rank * capacity + i] = v;
}

// after the pushes of this stream are complete: thread p publishes the epoch to peer p and waits for peer p's
__global__ void k_exchange_signal_wait(uint32_t world, uint32_t rank, uint32_t *const *__restrict__ peer_flags,
    uint32_t epoch, int *timed_out) {
    if (threadIdx.x >= world) return;
    publish(peer_flags[threadIdx.x] + rank, epoch);
    await(peer_flags[rank] + threadIdx.x, epoch, timed_out);
}

__global__ void k_exchange_sum(double *const *__restrict__ peer_recv, uint64_t count, uint64_t capacity, uint32_t world,
    uint32_t rank, const double *__restrict__ spans, uint64_t span_stride, uint64_t span_count, double *out) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double *recv = peer_recv[rank];
    double s = 0.0;
    for (uint32_t r = 0; r < world; r++) s += __ldcv(recv + (uint64_t) r * capacity + i);
    if (spans != nullptr) s /= span_of(spans, i, span_stride, span_count);
    out[i] = s;
}

size_t slab_bytes(const tskb_exchange &x) {
    return 2 * (size_t) x.world * x.capacity * sizeof(double) + 2 * (size_t) x.world * sizeof(uint32_t);
}
double *recv_of(const tskb_exchange &x, char *base, int parity) {
    return reinterpret_cast<double *>(base) + (size_t) parity * x.world * x.capacity;
}
uint32_t *flags_of(const tskb_exchange &x, char *base, int parity) {
    return reinterpret_cast<uint32_t *>(base + 2 * (size_t) x.world * x.capacity * sizeof(double)) + (size_t) parity * x.world;
}

int fail(const CudaFail &f) {
    last_error_string() = std::string(cudaGetErrorString(f.err)) + " at " + f.file + ":" + std::to_string(f.line);
    cudaGetLastError();
    return TSKB_ERR_CUDA;
}

void upload_tables(tskb_exchange &x) {
    std::vector<double *> pr(2 * x.world);
    std::vector<uint32_t *> pf(2 * x.world);
    for (int par = 0; par < 2; par++) {
        for (uint32_t p = 0; p < x.world; p++) {
            pr[par * x.world + p] = recv_of(x, x.peer_slab[p], par);
            pf[par * x.world + p] = flags_of(x, x.peer_slab[p], par);
        }
    }
    TSKB_CK(cudaMemcpy(x.d_peer_recv, pr.data(), pr.size() * sizeof(double *), cudaMemcpyHostToDevice));
    TSKB_CK(cudaMemcpy(x.d_peer_flags, pf.data(), pf.size() * sizeof(uint32_t *), cudaMemcpyHostToDevice));
}

}  // namespace
}  // namespace tskb

extern "C" {

using namespace tskb;

int tskb_exchange_create(int device, uint64_t capacity, uint32_t world, uint32_t rank, tskb_exchange_t **out) {
    if (out == nullptr || world == 0 || world > 64 || rank >= world || capacity == 0 || capacity > (1ull << 32)) {
        return TSKB_ERR_BAD_PARAM_VALUE;
    }
    *out = nullptr;
    tskb_exchange *x = new (std::nothrow) tskb_exchange();
    if (x == nullptr) return TSKB_ERR_NO_MEMORY;
    x->device = device;
    x->world = world;
    x->rank = rank;
    x->capacity = capacity;
    try {
        TSKB_CK(cudaSetDevice(device));
        TSKB_CK(cudaMalloc(&x->slab, slab_bytes(*x)));
        TSKB_CK(cudaMemset(x->slab, 0, slab_bytes(*x)));
        TSKB_CK(cudaMalloc(&x->d_peer_recv, 2 * world * sizeof(double *)));
        TSKB_CK(cudaMalloc(&x->d_peer_flags, 2 * world * sizeof(uint32_t *)));
        TSKB_CK(cudaHostAlloc(&x->h_timed_out, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
        *x->h_timed_out = 0;
        TSKB_CK(cudaStreamCreateWithFlags(&x->stream, cudaStreamNonBlocking));
        x->peer_slab.assign(world, nullptr);
        x->mapped.assign(world, false);
        x->peer_slab[rank] = x->slab;
        if (world == 1) {
            upload_tables(*x);
            x->connected = true;
        }
        TSKB_CK(cudaDeviceSynchronize());
    } catch (const CudaFail &f) {
        const int ret = f.err == cudaErrorMemoryAllocation ? TSKB_ERR_NO_MEMORY : fail(f);
        tskb_exchange_free(x);
        return ret;
    }
    *out = x;
    return 0;
}

int tskb_exchange_get_handle(const tskb_exchange_t *x, void *handle_out) {
    if (x == nullptr || handle_out == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    static_assert(sizeof(cudaIpcMemHandle_t) == TSKB_EXCHANGE_HANDLE_BYTES, "handle size");
    try {
        TSKB_CK(cudaSetDevice(x->device));
        cudaIpcMemHandle_t h;
        TSKB_CK(cudaIpcGetMemHandle(&h, x->slab));
        std::memcpy(handle_out, &h, sizeof(h));
        return 0;
    } catch (const CudaFail &f) {
        return fail(f);
    }
}

int tskb_exchange_connect(tskb_exchange_t *x, const void *handles) {
    if (x == nullptr || handles == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    std::lock_guard<std::mutex> lock(x->mu);
    if (x->connected) return TSKB_ERR_BAD_PARAM_VALUE;
    try {
        TSKB_CK(cudaSetDevice(x->device));
        for (uint32_t p = 0; p < x->world; p++) {
            if (p == x->rank) continue;
            cudaIpcMemHandle_t h;
            std::memcpy(&h, (const char *) handles + (size_t) p * sizeof(h), sizeof(h));
            void *base = nullptr;
            // opened from THIS rank's device: the mapping lives in its context, peer access is enabled with it
            TSKB_CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
            x->peer_slab[p] = (char *) base;
            x->mapped[p] = true;
        }
        upload_tables(*x);
        x->connected = true;
        return 0;
    } catch (const CudaFail &f) {
        return fail(f);
    }
}

int tskb_exchange_connect_local(tskb_exchange_t *x, tskb_exchange_t *const *members) {
    if (x == nullptr || members == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    std::lock_guard<std::mutex> lock(x->mu);
    if (x->connected) return TSKB_ERR_BAD_PARAM_VALUE;
    try {
        TSKB_CK(cudaSetDevice(x->device));
        for (uint32_t p = 0; p < x->world; p++) {
            const tskb_exchange *m = members[p];
            if (m == nullptr || m->world != x->world || m->capacity != x->capacity || m->rank != p) {
                return TSKB_ERR_BAD_PARAM_VALUE;
            }
            if (m->device != x->device) {
                int can = 0;
                TSKB_CK(cudaDeviceCanAccessPeer(&can, x->device, m->device));
                if (!can) {
                    last_error_string() = "no peer access between the two devices";
                    return TSKB_ERR_CUDA;
                }
                const cudaError_t e = cudaDeviceEnablePeerAccess(m->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) TSKB_CK(e);
                cudaGetLastError();
            }
            x->peer_slab[p] = m->slab;
        }
        upload_tables(*x);
        x->connected = true;
        return 0;
    } catch (const CudaFail &f) {
        return fail(f);
    }
}

int tskb_exchange_sum(tskb_exchange_t *x, const tskb_treeseq_t *engine, const double *d_local, uint64_t count,
    const double *d_spans, uint64_t span_stride, uint64_t span_count, double *d_out, uint32_t options) {
    if (x == nullptr || d_local == nullptr || d_out == nullptr || count == 0 || count > x->capacity
        || (d_spans != nullptr && (span_stride == 0 || span_count == 0))
        || (engine != nullptr && (engine->plan == nullptr || engine->plan->device != x->device))) {
        return TSKB_ERR_BAD_PARAM_VALUE;
    }
    std::lock_guard<std::mutex> lock(x->mu);
    if (!x->connected) return TSKB_ERR_BAD_PARAM_VALUE;
    // on the engine's stream the exchange follows the statistic that produced d_local without a host wait
    std::unique_lock<std::mutex> engine_lock;
    if (engine != nullptr) engine_lock = std::unique_lock<std::mutex>(engine->plan->mu);
    try {
        TSKB_CK(cudaSetDevice(x->device));
        cudaStream_t s = engine != nullptr ? engine->plan->stream : x->stream;
        const uint32_t epoch = ++x->epoch;
        const int par = epoch & 1;
        double *const *pr = x->d_peer_recv + par * x->world;
        uint32_t *const *pf = x->d_peer_flags + par * x->world;
        if (count <= FUSED_MAX) {
            k_exchange_fused<<<1, FUSED_THREADS, 0, s>>>(d_local, count, x->capacity, x->world, x->rank, pr, pf, epoch,
                d_spans, span_stride, span_count, d_out, x->h_timed_out);
        } else {
            k_exchange_push<<<grid_for(count, TBX), TBX, 0, s>>>(d_local, count, x->capacity, x->world, x->rank, pr);
            k_exchange_signal_wait<<<1, 64, 0, s>>>(x->world, x->rank, pf, epoch, x->h_timed_out);
            k_exchange_sum<<<grid_for(count, TBX), TBX, 0, s>>>(pr, count, x->capacity, x->world, x->rank, d_spans,
                span_stride, span_count, d_out);
        }
        TSKB_CK_LAUNCH();
        if (options & TSKB_EXCHANGE_ASYNC) return 0;
        TSKB_CK(cudaStreamSynchronize(s));
        if (*(volatile int *) x->h_timed_out) {
            last_error_string() = "exchange: a peer's partial did not arrive";
            return TSKB_ERR_CUDA;
        }
        return 0;
    } catch (const CudaFail &f) {
        return fail(f);
    }
}

int tskb_exchange_status(tskb_exchange_t *x, const tskb_treeseq_t *engine) {
    if (x == nullptr) return TSKB_ERR_BAD_PARAM_VALUE;
    std::lock_guard<std::mutex> lock(x->mu);
    try {
        TSKB_CK(cudaSetDevice(x->device));
        TSKB_CK(cudaStreamSynchronize(engine != nullptr && engine->plan != nullptr ? engine->plan->stream : x->stream));
        if (*(volatile int *) x->h_timed_out) {
            last_error_string() = "exchange: a peer's partial did not arrive";
            return TSKB_ERR_CUDA;
        }
        return 0;
    } catch (const CudaFail &f) {
        return fail(f);
    }
}

int tskb_exchange_free(tskb_exchange_t *x) {
    if (x == nullptr) return 0;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (size_t p = 0; p < x->peer_slab.size(); p++) {
        if (x->mapped[p] && x->peer_slab[p] != nullptr) cudaIpcCloseMemHandle(x->peer_slab[p]);
    }
    if (x->slab) cudaFree(x->slab);
    if (x->d_peer_recv) cudaFree(x->d_peer_recv);
    if (x->d_peer_flags) cudaFree(x->d_peer_flags);
    if (x->h_timed_out) cudaFreeHost(x->h_timed_out);
    if (x->stream) cudaStreamDestroy(x->stream);
    cudaGetLastError();
    delete x;
    return 0;
}

}  // extern "C"
