// Micro-benchmark behind the scatter analysis in DESIGN.md §5.2: n threads x 3 windows of uniformly random bucket ids
// (3 x 2^19 buckets, ~128 entries per bucket — the geometry of one window group of the 2^26 MSM) doing
//   red   : atomicAdd without a returned value (what the histogram pass does)
//   atom  : atomicAdd with the returned value consumed in a register
//   atom+st: returned position used for a dependent 4-byte store (what the scatter pass does)
//   st    : the same scattered 4-byte stores without any atomic
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scatter_bench scatter_bench.cu     (tool, not part of the library)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return (uint32_t)x;
}
template <int MODE> __global__ void k(size_t n, uint32_t nb, uint32_t cap, uint32_t *cursor, uint32_t *sorted, uint32_t *sink) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t acc = 0;
    for (int w = 0; w < 3; w++) {
        const uint32_t b = (uint32_t)w * nb + (mix(i * 3 + w) & (nb - 1));
        if (MODE == 0) atomicAdd(&cursor[b], 1u);
        else if (MODE == 1) acc += atomicAdd(&cursor[b], 1u);
        else if (MODE == 2) { uint32_t pos = atomicAdd(&cursor[b], 1u); sorted[(size_t)b * cap + (pos & (cap - 1))] = (uint32_t)i; }
        else sorted[(size_t)b * cap + ((uint32_t)(i >> 19) & (cap - 1))] = (uint32_t)i;
    }
    if (MODE == 1 && acc == 0xffffffffu) *sink = acc;
}
int main() {
    const size_t n = (size_t)1 << 26;
    const uint32_t nb = 1u << 19, cap = 256;
    uint32_t *cursor, *sorted, *sink;
    cudaMalloc(&cursor, 3 * (size_t)nb * 4);
    cudaMalloc(&sorted, 3 * (size_t)nb * cap * 4);
    cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[4] = {"red", "atom", "atom+store", "store"};
    for (int mode = 0; mode < 4; mode++) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; rep++) {
            cudaMemset(cursor, 0, 3 * (size_t)nb * 4);
            cudaEventRecord(e0);
            const unsigned blocks = (unsigned)((n + 255) / 256);
            if (mode == 0) k<0><<<blocks, 256>>>(n, nb, cap, cursor, sorted, sink);
            else if (mode == 1) k<1><<<blocks, 256>>>(n, nb, cap, cursor, sorted, sink);
            else if (mode == 2) k<2><<<blocks, 256>>>(n, nb, cap, cursor, sorted, sink);
            else k<3><<<blocks, 256>>>(n, nb, cap, cursor, sorted, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        printf("{\"mode\": \"%s\", \"entries\": %zu, \"ms\": %.3f, \"G_entries_per_s\": %.1f}\n", names[mode], 3 * n, best, 3.0 * n / best / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
