// microbench_gather.cu -- ceiling for the propagation kernel's addend gather on this GPU:
// random 4 / 8-byte gathers (one per 32-byte sector) from a table of a given size, indices
// streamed coalesced, results written coalesced.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
template <typename T, int IPT>
__global__ void k_gather(const uint32_t *__restrict__ idx, const T *tab, T *out, size_t n) {
    size_t base = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) * IPT;
    if (base + IPT > n) return;
    uint32_t ix[IPT];
    const uint4 *q = reinterpret_cast<const uint4 *>(idx + base);
#pragma unroll
    for (int i = 0; i < IPT / 4; i++) { uint4 t = q[i]; ix[4*i]=t.x; ix[4*i+1]=t.y; ix[4*i+2]=t.z; ix[4*i+3]=t.w; }
    T v[IPT];
#pragma unroll
    for (int i = 0; i < IPT; i++) v[i] = __ldcg(tab + ix[i]);
#pragma unroll
    for (int i = 0; i < IPT; i++) out[base + i] = v[i];
}
__global__ void k_fill(uint32_t *idx, size_t n, uint32_t mod, uint32_t local) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t x = i * 0x9E3779B97F4A7C15ull + 12345; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    uint32_t r = (uint32_t) (x % mod);
    if (local) {  // locality model: target within a window of `local` entries sliding with i
        uint64_t centre = (uint64_t) ((double) i / (double) n * (double) mod);
        r = (uint32_t) ((centre + (x % local)) % mod);
    }
    idx[i] = r;
}
template <typename T>
void run(size_t n, size_t tab_entries, uint32_t local) {
    uint32_t *idx; T *tab, *out;
    cudaMalloc(&idx, n * 4); cudaMalloc(&tab, tab_entries * sizeof(T)); cudaMalloc(&out, n * sizeof(T));
    cudaMemset(tab, 1, tab_entries * sizeof(T));
    k_fill<<<(n + 255) / 256, 256>>>(idx, n, (uint32_t) tab_entries, local);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int it = 0; it < 5; it++) {
        cudaEventRecord(a);
        k_gather<T, 8><<<(n / 8 + 255) / 256, 256>>>(idx, tab, out, n);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    printf("elem %zuB table %7.1f MB local %9u: %8.3f ms  %7.2f G gathers/s\n", sizeof(T), tab_entries * sizeof(T) / 1e6, local, best, n / best / 1e6);
    cudaFree(idx); cudaFree(tab); cudaFree(out);
}
int main() {
    size_t n = 80u << 20;
    for (size_t mb : {16, 48, 96, 200, 400, 800}) run<int>(n, mb * 1000000 / 4, 0);
    for (size_t mb : {200, 400}) run<int2>(n, mb * 1000000 / 8, 0);
    for (uint32_t local : {1u << 16, 1u << 20, 1u << 22, 1u << 24}) run<int>(n, 50000000, local);
    return 0;
}
