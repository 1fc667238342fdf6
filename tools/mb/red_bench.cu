// microbenchmark: throughput of global fp64 reductions (REDG.E.ADD.F64) by address distribution
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_red(double *bins, uint32_t nbins, uint32_t per_thread, uint32_t mode, uint32_t stride) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s = t * 2654435761u + 12345u;
    for (uint32_t i = 0; i < per_thread; i++) {
        uint32_t a;
        if (mode == 0) { s = s * 1664525u + 1013904223u; a = (s >> 8) % nbins; }           // random
        else if (mode == 1) { a = (uint32_t) (((uint64_t) t * per_thread + i) * stride % nbins); } // increasing per thread
        else { a = (blockIdx.x * 7u + (i >> 2)) % nbins; }                                    // warp-uniform
        atomicAdd(&bins[a], 1.0);
    }
}
int main() {
    double *bins; size_t maxb = 1u << 24;
    cudaMalloc(&bins, maxb * 8); cudaMemset(bins, 0, maxb * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const uint32_t grid = 148 * 8, tb = 256, per = 256;
    const double total = (double) grid * tb * per;
    struct { const char *name; uint32_t nbins, mode, stride; } cases[] = {
        {"random over 2000 bins", 2000, 0, 1}, {"random over 64k bins", 65536, 0, 1},
        {"random over 4.4M bins", 4400000, 0, 1}, {"random over 16M bins", 1u << 24, 0, 1},
        {"per-thread increasing, stride 97, 4.4M", 4400000, 1, 97}, {"warp-uniform 2000", 2000, 2, 1}};
    for (auto &c : cases) {
        k_red<<<grid, tb>>>(bins, c.nbins, 8, c.mode, c.stride);
        cudaEventRecord(e0);
        k_red<<<grid, tb>>>(bins, c.nbins, per, c.mode, c.stride);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-45s %8.3f ms  %7.2f G red/s\n", c.name, ms, total / ms / 1e6);
    }
    return 0;
}
