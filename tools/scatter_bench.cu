// Micro-benchmark behind the scatter analysis in DESIGN.md §5.2: n threads x {1, 2, 3, 6} windows of uniformly random bucket ids
// (2^19 buckets per window, ~128 entries per bucket — the geometry of the 2^26 MSM; 3 windows = one window group) doing
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
template <int MODE> __global__ void k(size_t n, uint32_t nb, uint32_t cap, uint32_t *cursor, uint32_t *sorted, uint32_t *sink, int windows) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t acc = 0;
    for (int w = 0; w < windows; w++) {
        const uint32_t b = (uint32_t)w * nb + (mix(i * 8 + w) & (nb - 1));
        if (MODE == 0) atomicAdd(&cursor[b], 1u);
        else if (MODE == 1) acc += atomicAdd(&cursor[b], 1u);
        else if (MODE == 2) { uint32_t pos = atomicAdd(&cursor[b], 1u); sorted[(size_t)b * cap + (pos & (cap - 1))] = (uint32_t)i; }
        else sorted[(size_t)b * cap + ((uint32_t)(i >> 19) & (cap - 1))] = (uint32_t)i;
    }
    if (MODE == 1 && acc == 0xffffffffu) *sink = acc;
}
int main(int argc, char **argv) {
    const size_t n = (size_t)1 << 26;
    const uint32_t nb = 1u << 19, cap = 256;
    const int maxw = 6;
    uint32_t *cursor, *sorted, *sink;
    cudaMalloc(&cursor, maxw * (size_t)nb * 4);
    cudaMalloc(&sorted, maxw * (size_t)nb * cap * 4);
    cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[4] = {"red", "atom", "atom+store", "store"};
    // windows per pass = open write fronts: 1 -> 16 MiB of 32-byte sectors, 3 -> 48 MiB (the library's group size), 6 -> 96 MiB
    for (int windows : {1, 2, 3, 6})
    for (int mode = 0; mode < 4; mode++) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; rep++) {
            cudaMemset(cursor, 0, maxw * (size_t)nb * 4);
            cudaEventRecord(e0);
            const unsigned blocks = (unsigned)((n + 255) / 256);
            if (mode == 0) k<0><<<blocks, 256>>>(n, nb, cap, cursor, sorted, sink, windows);
            else if (mode == 1) k<1><<<blocks, 256>>>(n, nb, cap, cursor, sorted, sink, windows);
            else if (mode == 2) k<2><<<blocks, 256>>>(n, nb, cap, cursor, sorted, sink, windows);
            else k<3><<<blocks, 256>>>(n, nb, cap, cursor, sorted, sink, windows);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        printf("{\"windows\": %d, \"mode\": \"%s\", \"entries\": %zu, \"ms\": %.3f, \"G_entries_per_s\": %.1f}\n", windows, names[mode], windows * n, best, (double)windows * n / best / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
