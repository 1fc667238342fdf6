// microbenchmark: one scattered reduction per streamed item (20 B per item, coalesced) into a
// T-entry array, by reduction type: fp64, u64, f32, u32.  Is the L2's fp64 add the slow one?
// (The branch summary issues ~1.2 fp64 reductions per piece to 2.75 M addresses on the C2 ARG.)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <typename T>
__global__ void k(const int *a, const double *b, const uint32_t *c0, T *D, uint32_t n) {
    const uint32_t tile = 4 * blockDim.x;
    for (uint32_t base = blockIdx.x * tile; base + tile <= n; base += gridDim.x * tile) {
        int s[4]; double bl[4]; uint32_t p0[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t j = base + q * blockDim.x + threadIdx.x;
            s[q] = a[j]; bl[q] = b[j]; p0[q] = c0[j];
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            double x = (double) s[q];
            double G = bl[q] * (x * (100000.0 - x) * 1e-10) + 1.0;
            atomicAdd(D + p0[q], (T) G);
        }
    }
}
template <typename T>
void run(const char *name, const int *a, const double *b, const uint32_t *c0, void *D, uint32_t n, uint32_t T_) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int bps : {4, 8, 16}) {
        float best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            k<T><<<148 * bps, 256>>>(a, b, c0, (T *) D, n);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        printf("%-4s addresses %8u blocks/SM %2d : %7.3f ms  %6.1f G red/s\n", name, T_, bps, best, (double) n / best / 1e6);
    }
}
int main() {
    const uint32_t n = 48u << 20;
    int *a; double *b; void *D; uint32_t *c0;
    cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 8); cudaMalloc(&c0, n * 4); cudaMalloc(&D, 4400000 * 8);
    cudaMemset(a, 1, n * 4); cudaMemset(b, 0, n * 8);
    uint32_t *h = (uint32_t *) malloc(n * 4);
    for (uint32_t T : {2750000u, 128000u, 2000u}) {
        uint32_t s = 12345;
        for (uint32_t i = 0; i < n; i++) { s = s * 1664525u + 1013904223u; h[i] = (s >> 7) % T; }
        cudaMemcpy(c0, h, n * 4, cudaMemcpyHostToDevice);
        cudaMemset(D, 0, 4400000 * 8);
        run<double>("f64", a, b, c0, D, n, T);
        run<unsigned long long>("u64", a, b, c0, D, n, T);
        run<float>("f32", a, b, c0, D, n, T);
        run<unsigned int>("u32", a, b, c0, D, n, T);
    }
    return 0;
}
