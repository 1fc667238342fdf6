// exchange.cu -- summing the ranks' per-window partials over NVLink peer memory, without NCCL.
//
// The only exchange on the sharded path is the sum of W x M doubles per statistic (a few KB).  Through
// NCCL that costs a host-side launch per call plus the collective's own latency, more than the transfer
// itself.  Here every rank PUSHES its partial into slot [rank] of a receive buffer on every peer
// (plain stores over NVLink into memory the peers exported by CUDA IPC), publishes an epoch number in
// the peer's flag array after a system-scope fence, and each rank then waits on its own flags (device
// side, bounded) and adds the world's slots in rank order -- the same order on every rank, so the sum
// is bit-identical everywhere -- and span-normalises (trees.c:1920-1934).  Buffers are double-buffered
// by epoch parity: a rank can be at most one step ahead of a peer, because finishing a step needs
// every peer's partial of that step.
#include <cuda_runtime.h>
#include <stdint.h>

#include "plan.cuh"

namespace tskb {
namespace {

constexpr int TBX = 256;
constexpr uint32_t WAIT_SPINS = 1u << 24;  // ~ seconds: a peer that died must not hang this GPU

__global__ void k_exchange_push(const double *__restrict__ local, uint64_t count, uint32_t world, uint32_t rank,
    double *const *__restrict__ peer_recv) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double v = local[i];
    for (uint32_t p = 0; p < world; p++) peer_recv[p][(uint64_t) rank * count + i] = v;
}

// after the pushes of this stream are complete: one thread per peer publishes the epoch
__global__ void k_exchange_signal(uint32_t world, uint32_t rank, uint32_t *const *__restrict__ peer_flags, uint32_t epoch) {
    const uint32_t p = threadIdx.x;
    if (p >= world) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flags[p] + rank), "r"(epoch) : "memory");
}

__global__ void k_exchange_wait(const uint32_t *flags, uint32_t world, uint32_t epoch, int *timed_out) {
    const uint32_t p = threadIdx.x;
    if (p >= world) return;
    uint32_t v = 0, spins = 0;
    while (true) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + p) : "memory");
        if ((int32_t) (v - epoch) >= 0) break;
        if (++spins > WAIT_SPINS) {
            *timed_out = 1;
            break;
        }
        __nanosleep(100);
    }
}

// out[i] = sum over ranks (in rank order) of recv[r][i], divided by the span of the element's window
__global__ void k_exchange_sum(const double *__restrict__ recv, uint64_t count, uint32_t world,
    const double *__restrict__ spans, uint64_t span_stride, uint64_t span_count, double *out) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double s = 0.0;
    for (uint32_t r = 0; r < world; r++) s += __ldcv(recv + (uint64_t) r * count + i);  // written by peers: no stale cache lines
    if (spans != nullptr) s /= spans[(i / span_stride) % span_count];
    out[i] = s;
}

}  // namespace
}  // namespace tskb

extern "C" {

using namespace tskb;

int tskb_enable_peer_access(int device, int peer_device) {
    if (device == peer_device) return 0;
    int can = 0;
    if (cudaSetDevice(device) != cudaSuccess || cudaDeviceCanAccessPeer(&can, device, peer_device) != cudaSuccess || !can) {
        cudaGetLastError();
        last_error_string() = "no peer access between the two devices";
        return TSKB_ERR_CUDA;
    }
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        last_error_string() = cudaGetErrorString(e);
        cudaGetLastError();
        return TSKB_ERR_CUDA;
    }
    cudaGetLastError();
    return 0;
}

int tskb_exchange_sum(const tskb_treeseq_t *self, const double *d_local, uint64_t count, uint32_t world,
    uint32_t rank, double *const *peer_recv, uint32_t *const *peer_flags, const double *d_recv,
    const uint32_t *d_flags, uint32_t epoch, const double *d_spans, uint64_t span_stride, uint64_t span_count,
    double *d_out) {
    if (self == nullptr || self->plan == nullptr || d_local == nullptr || d_out == nullptr || peer_recv == nullptr
        || peer_flags == nullptr || d_recv == nullptr || d_flags == nullptr || world == 0 || rank >= world
        || world > 64 || (d_spans != nullptr && (span_stride == 0 || span_count == 0))) {
        return TSKB_ERR_BAD_PARAM_VALUE;
    }
    const Plan &P = *self->plan;
    std::lock_guard<std::mutex> lock(P.mu);
    try {
        TSKB_CK(cudaSetDevice(P.device));
        cudaStream_t s = P.stream;
        Arena &A = P.arena;
        A.reset();
        // the peers' buffer addresses for this call (small tables in device memory)
        double **d_pr = A.get<double *>(world);
        uint32_t **d_pf = A.get<uint32_t *>(world);
        int *d_to = A.get<int>(1);
        TSKB_CK(cudaMemcpyAsync(d_pr, peer_recv, world * sizeof(double *), cudaMemcpyHostToDevice, s));
        TSKB_CK(cudaMemcpyAsync(d_pf, peer_flags, world * sizeof(uint32_t *), cudaMemcpyHostToDevice, s));
        TSKB_CK(cudaMemsetAsync(d_to, 0, sizeof(int), s));
        if (count) {
            k_exchange_push<<<grid_for(count, TBX), TBX, 0, s>>>(d_local, count, world, rank, d_pr);
            TSKB_CK_LAUNCH();
        }
        k_exchange_signal<<<1, 64, 0, s>>>(world, rank, d_pf, epoch);
        k_exchange_wait<<<1, 64, 0, s>>>(d_flags, world, epoch, d_to);
        if (count) {
            k_exchange_sum<<<grid_for(count, TBX), TBX, 0, s>>>(d_recv, count, world, d_spans, span_stride, span_count, d_out);
        }
        TSKB_CK_LAUNCH();
        int h_to = 0;
        TSKB_CK(cudaMemcpyAsync(&h_to, d_to, sizeof(int), cudaMemcpyDeviceToHost, s));
        TSKB_CK(cudaStreamSynchronize(s));
        if (h_to) {
            last_error_string() = "exchange: a peer's partial did not arrive";
            return TSKB_ERR_CUDA;
        }
        return 0;
    } catch (const CudaFail &f) {
        last_error_string() = std::string(cudaGetErrorString(f.err)) + " at " + f.file + ":" + std::to_string(f.line);
        cudaGetLastError();
        return TSKB_ERR_CUDA;
    }
}

}  // extern "C"
