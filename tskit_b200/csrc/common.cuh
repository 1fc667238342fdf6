// common.cuh -- small host/device utilities shared by the engine's translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>
#include <vector>

namespace tskb {

struct CudaFail {
    cudaError_t err;
    const char *what;
    const char *file;
    int line;
};

std::string &last_error_string();

#define TSKB_CK(expr)                                                                   \
    do {                                                                                \
        cudaError_t e__ = (expr);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            throw ::tskb::CudaFail{ e__, #expr, __FILE__, __LINE__ };                   \
        }                                                                               \
    } while (0)

#define TSKB_CK_LAUNCH() TSKB_CK(cudaGetLastError())

// While a plan is being built its many temporaries come from the device's stream-ordered memory pool
// (cudaMallocAsync / cudaFreeAsync on the build stream): a freed block is handed to the next
// allocation without a trip to the driver, and nothing synchronises the device.  Outside a build the
// arrays fall back to cudaMalloc / cudaFree.
inline cudaStream_t &alloc_stream() {
    static thread_local cudaStream_t s = nullptr;
    return s;
}

// Owning device array; sized once at plan build.
template <typename T>
struct DevArray {
    T *p = nullptr;
    size_t n = 0;
    DevArray() = default;
    DevArray(const DevArray &) = delete;
    DevArray &operator=(const DevArray &) = delete;
    DevArray(DevArray &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevArray &operator=(DevArray &&o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevArray() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (count > 0) {
            if (alloc_stream() != nullptr) {
                TSKB_CK(cudaMallocAsync((void **) &p, count * sizeof(T), alloc_stream()));
            } else {
                TSKB_CK(cudaMalloc(&p, count * sizeof(T)));
            }
        }
    }
    void release() {
        if (p != nullptr) {
            if (alloc_stream() != nullptr) {
                cudaFreeAsync(p, alloc_stream());
            } else {
                cudaFree(p);
            }
            p = nullptr;
        }
        n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
    void upload(const T *host, size_t count, cudaStream_t s) {
        alloc(count);
        if (count > 0) {
            TSKB_CK(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
        }
    }
    std::vector<T> download(cudaStream_t s) const {
        std::vector<T> out(n);
        if (n > 0) {
            TSKB_CK(cudaMemcpyAsync(out.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost, s));
            TSKB_CK(cudaStreamSynchronize(s));
        }
        return out;
    }
};

// Bump allocator over one growable device slab: per-call scratch without
// cudaMalloc on the hot path.  Not thread safe; guarded by the plan's mutex.
struct Arena {
    char *base = nullptr;
    size_t cap = 0;
    size_t off = 0;
    size_t high = 0;
    std::vector<void *> overflow;  // blocks malloc'd when the slab was too small
    ~Arena() { destroy(); }
    void destroy() {
        for (void *q : overflow) cudaFree(q);
        overflow.clear();
        if (base) cudaFree(base);
        base = nullptr; cap = 0; off = 0;
    }
    void reset() {
        // grow the slab to the high-water mark of the previous call
        if (!overflow.empty()) {
            cudaDeviceSynchronize();
            for (void *q : overflow) cudaFree(q);
            overflow.clear();
            if (base) cudaFree(base);
            base = nullptr; cap = 0;
        }
        off = 0;
        if (high > cap) {
            if (base) { cudaDeviceSynchronize(); cudaFree(base); base = nullptr; }
            cap = 0;
            const size_t want = high + (high >> 3) + (1 << 20);
            high = 0;
            char *q = nullptr;
            TSKB_CK(cudaMalloc(&q, want));  // on failure the arena stays empty and consistent
            base = q;
            cap = want;
        }
        high = 0;
    }
    template <typename T>
    T *get(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
        if (bytes == 0) bytes = 256;
        high += bytes;
        if (off + bytes <= cap) {
            T *r = reinterpret_cast<T *>(base + off);
            off += bytes;
            return r;
        }
        void *q = nullptr;
        TSKB_CK(cudaMalloc(&q, bytes));
        overflow.push_back(q);
        return reinterpret_cast<T *>(q);
    }
};

inline unsigned ceil_log2(uint64_t x) {
    unsigned b = 0;
    while ((uint64_t(1) << b) < x) b++;
    return b;
}

inline int grid_for(size_t n, int block) {
    size_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    return (int) g;
}

// ---- device helpers ----
// first index in [0, n) with a[idx] >= x
template <typename T>
__host__ __device__ inline uint32_t lower_bound_dev(const T *a, uint32_t n, T x) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// first index in [0, n) with a[idx] > x
template <typename T>
__host__ __device__ inline uint32_t upper_bound_dev(const T *a, uint32_t n, T x) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

}  // namespace tskb
