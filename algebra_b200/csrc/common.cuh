// common.cuh — error plumbing, launch accounting and small host helpers shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/algebra_b200.h"
#include "fp.cuh"

namespace ab200 {

void set_last_error(const std::string &msg);
// once per device: keep freed stream-ordered allocations in the pool (default release threshold 0 gives the memory back at
// every synchronisation, and re-allocating GBs per call costs tens of ms)
int ensure_device_init();
extern std::atomic<unsigned long long> g_launches;

inline int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    set_last_error(buf);
    return (int)e;
}
#define AB_CUDA(expr)                                                             \
    do {                                                                          \
        cudaError_t _e = (expr);                                                  \
        if (_e != cudaSuccess) return ::ab200::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)
// after a kernel launch: count it and surface launch-configuration errors
#define AB_LAUNCHED()                                                             \
    do {                                                                          \
        ::ab200::g_launches.fetch_add(1, std::memory_order_relaxed);              \
        cudaError_t _e = cudaGetLastError();                                      \
        if (_e != cudaSuccess) return ::ab200::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

// 8 or 12 limbs passed by value to kernels
template <int L> struct LimbArg {
    uint32_t v[L];
};

// loads/stores of one field element as 16-byte vectors (L = 8 -> 2 x uint4, L = 12 -> 3 x uint4)
template <int L> __device__ __forceinline__ void load_limbs(uint32_t *r, const uint32_t *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
#pragma unroll
    for (int i = 0; i < L / 4; i++) {
        uint4 t = q[i];
        r[4 * i] = t.x; r[4 * i + 1] = t.y; r[4 * i + 2] = t.z; r[4 * i + 3] = t.w;
    }
}
template <int L> __device__ __forceinline__ void load_limbs_nc(uint32_t *r, const uint32_t *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
#pragma unroll
    for (int i = 0; i < L / 4; i++) {
        uint4 t = __ldg(q + i);
        r[4 * i] = t.x; r[4 * i + 1] = t.y; r[4 * i + 2] = t.z; r[4 * i + 3] = t.w;
    }
}
template <int L> __device__ __forceinline__ void store_limbs(uint32_t *p, const uint32_t *r) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
#pragma unroll
    for (int i = 0; i < L / 4; i++) q[i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
}

}  // namespace ab200
