// microbenchmark: stream 20 B per item (coalesced) and issue R reductions per item to pseudo-random
// slots of a 4.4M-entry fp64 array -- the shape of the branch summary.  Variants: items per thread,
// blocks per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int IPT>
__global__ void k(const int *a, const double *b, const uint32_t *c0, const uint32_t *c1, double *D, uint32_t n,
    int reds) {
    const uint32_t tile = IPT * blockDim.x;
    for (uint32_t base = blockIdx.x * tile; base + tile <= n; base += gridDim.x * tile) {
        int s[IPT]; double bl[IPT]; uint32_t p0[IPT], p1[IPT];
#pragma unroll
        for (int q = 0; q < IPT; q++) {
            uint32_t j = base + q * blockDim.x + threadIdx.x;
            s[q] = a[j]; bl[q] = b[j]; p0[q] = c0[j]; p1[q] = c1[j];
        }
#pragma unroll
        for (int q = 0; q < IPT; q++) {
            double x = (double) s[q];
            double G = bl[q] * (x * (100000.0 - x) * 1e-10 + (100000.0 - x) * x * 1e-10);
            if (reds >= 1) atomicAdd(D + p0[q], G);
            if (reds >= 2) atomicAdd(D + p1[q], -G);
        }
    }
}
int main() {
    const uint32_t n = 48u << 20, T = 4400000;
    int *a; double *b, *D; uint32_t *c0, *c1;
    cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 8); cudaMalloc(&c0, n * 4); cudaMalloc(&c1, n * 4); cudaMalloc(&D, T * 8);
    cudaMemset(a, 1, n * 4); cudaMemset(b, 0, n * 8); cudaMemset(D, 0, T * 8);
    uint32_t *h = (uint32_t *) malloc(n * 4);
    uint32_t s = 12345;
    for (uint32_t i = 0; i < n; i++) { s = s * 1664525u + 1013904223u; h[i] = (s >> 7) % T; }
    cudaMemcpy(c0, h, n * 4, cudaMemcpyHostToDevice);
    for (uint32_t i = 0; i < n; i++) { s = s * 1664525u + 1013904223u; h[i] = (s >> 7) % T; }
    cudaMemcpy(c1, h, n * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int reds = 0; reds <= 2; reds++) {
        for (int bps : {2, 4, 8}) {
            for (int ipt : {4, 8}) {
                float best = 1e9;
                for (int rep = 0; rep < 3; rep++) {
                    cudaEventRecord(e0);
                    if (ipt == 4) k<4><<<148 * bps, 256>>>(a, b, c0, c1, D, n, reds);
                    else k<8><<<148 * bps, 256>>>(a, b, c0, c1, D, n, reds);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    if (ms < best) best = ms;
                }
                printf("reds/item %d  blocks/SM %d  items/thread %d : %7.3f ms  (%5.2f TB/s streamed, %6.1f G red/s)\n",
                    reds, bps, ipt, best, n * 20.0 / best / 1e9, reds * (double) n / best / 1e6);
            }
        }
    }
    return 0;
}
