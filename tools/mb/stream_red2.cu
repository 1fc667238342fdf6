// microbenchmark 2: the branch-summary shape with (a) half-populated reduction instructions and
// (b) clustered addresses: does the cost of fp64 reductions depend on lanes per instruction or on
// address locality?  1 reduction per item on average in every variant.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(const int *a, const double *b, const uint32_t *c0, const uint32_t *c1, double *D, uint32_t n) {
    const uint32_t tile = 4 * blockDim.x;
    for (uint32_t base = blockIdx.x * tile; base + tile <= n; base += gridDim.x * tile) {
        int s[4]; double bl[4]; uint32_t p0[4], p1[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t j = base + q * blockDim.x + threadIdx.x;
            s[q] = a[j]; bl[q] = b[j]; p0[q] = c0[j]; p1[q] = c1[j];
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            double x = (double) s[q];
            double G = bl[q] * (x * (100000.0 - x) * 1e-10 + (100000.0 - x) * x * 1e-10);
            if (MODE == 0) { atomicAdd(D + p0[q], G); }                       // one full instruction
            if (MODE == 1) { if (p0[q] & 1) atomicAdd(D + p0[q], G); else atomicAdd(D + p1[q], -G); }  // two half-full
            if (MODE == 2) { if (p0[q] & 1) atomicAdd(D + p0[q], G); if (p1[q] & 1) atomicAdd(D + p1[q], -G); } // two independent halves
        }
    }
}
int main() {
    const uint32_t n = 48u << 20, T = 4400000;
    int *a; double *b, *D; uint32_t *c0, *c1;
    cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 8); cudaMalloc(&c0, n * 4); cudaMalloc(&c1, n * 4); cudaMalloc(&D, T * 8);
    cudaMemset(a, 1, n * 4); cudaMemset(b, 0x3f, n * 8); cudaMemset(D, 0, T * 8);  // b = 0x3f3f... = 4.7e-4 (non-zero addends)
    uint32_t *h0 = (uint32_t *) malloc(n * 4), *h1 = (uint32_t *) malloc(n * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int dist = 0; dist < 3; dist++) {
        uint32_t s = 12345, cur = 0;
        for (uint32_t i = 0; i < n; i++) {
            s = s * 1664525u + 1013904223u;
            if (dist == 0) { h0[i] = (s >> 7) % T; }                                   // random
            else if (dist == 1) { if ((i & 31) == 0) cur = (s >> 7) % T; cur = (cur + 1 + ((s >> 3) & 7)) % T; h0[i] = cur; }  // runs of 32 with steps 1..8
            else { if ((i & 31) == 0) cur = (s >> 7) % T; cur = (cur + 1 + ((s >> 3) & 1023)) % T; h0[i] = cur; }              // runs with steps 1..1024
            s = s * 1664525u + 1013904223u; h1[i] = (s >> 7) % T;
        }
        cudaMemcpy(c0, h0, n * 4, cudaMemcpyHostToDevice); cudaMemcpy(c1, h1, n * 4, cudaMemcpyHostToDevice);
        const char *dn[] = {"random", "runs of 32, steps 1-8", "runs of 32, steps 1-1024"};
        for (int mode = 0; mode < 3; mode++) {
            float best = 1e9;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<148 * 4, 256>>>(a, b, c0, c1, D, n);
                if (mode == 1) k<1><<<148 * 4, 256>>>(a, b, c0, c1, D, n);
                if (mode == 2) k<2><<<148 * 4, 256>>>(a, b, c0, c1, D, n);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("addresses %-26s mode %d : %7.3f ms (%6.1f G red/s)\n", dn[dist], mode, best, (double) n / best / 1e6);
        }
    }
    return 0;
}
