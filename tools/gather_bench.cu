// Random-gather microbenchmark for the MSM's level-1 operand fetches: every thread reads R bytes (as 16-byte loads) from the start of a
// random record of a table of records with stride S bytes (table >> L2), ILP independent gathers in flight per thread.  Reports
// G gathers/s and useful GB/s per (S, R): decides whether padding the 96-byte affine points to 128 bytes / a compact 64-byte x-plane
// pays (fewer 128-byte lines per gather).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
template <int R16, int ILP> __global__ void __launch_bounds__(256) gather(const uint4 *tab, uint32_t nrec_mask, uint32_t stride16, int iters, uint32_t *out, uint32_t seed) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
    for (int it = 0; it < iters; it++) {
        uint4 v[ILP][R16];
#pragma unroll
        for (int k = 0; k < ILP; k++) {
            const uint32_t rec = hash32(t * 2654435761u + (it * ILP + k) * 40503u + seed) & nrec_mask;
            const uint4 *p = tab + (size_t)rec * stride16;
#pragma unroll
            for (int j = 0; j < R16; j++) v[k][j] = __ldg(p + j);
        }
#pragma unroll
        for (int k = 0; k < ILP; k++)
#pragma unroll
            for (int j = 0; j < R16; j++) acc ^= v[k][j].x ^ v[k][j].y ^ v[k][j].z ^ v[k][j].w;
    }
    out[t] = acc;
}
template <int R16> static void run(const uint4 *tab, size_t bytes, int stride, uint32_t *out, int sms) {
    const uint32_t stride16 = stride / 16;
    uint32_t nrec = 1; while ((size_t)nrec * 2 * stride <= bytes) nrec *= 2;
    const int blocks = sms * 8, iters = 64;
    constexpr int ILP = 4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather<R16, ILP><<<blocks, 256>>>(tab, nrec - 1, stride16, 4, out, 1); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); gather<R16, ILP><<<blocks, 256>>>(tab, nrec - 1, stride16, iters, out, 7 + r); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double gathers = (double)blocks * 256 * iters * ILP;
    printf("{\"stride\":%d,\"read_bytes\":%d,\"records\":%u,\"ms\":%.3f,\"G_gathers_per_s\":%.2f,\"useful_GB_per_s\":%.1f}\n", stride, R16 * 16, nrec, best,
           gathers / best / 1e6, gathers * R16 * 16 / best / 1e6);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const size_t bytes = (size_t)4 << 30;
    uint4 *tab; uint32_t *out;
    cudaMalloc(&tab, bytes); cudaMemset(tab, 1, bytes); cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    printf("{\"device\":\"%s\",\"table_GiB\":4}\n", p.name);
    run<3>(tab, bytes, 96, out, p.multiProcessorCount);    // x of an AoS affine point (48 of 96 bytes)
    run<6>(tab, bytes, 96, out, p.multiProcessorCount);    // whole AoS affine point
    run<3>(tab, bytes, 128, out, p.multiProcessorCount);   // x of a point padded to one 128-byte line
    run<6>(tab, bytes, 128, out, p.multiProcessorCount);   // whole point padded to one line
    run<3>(tab, bytes, 64, out, p.multiProcessorCount);    // compact x-plane, 64-byte records
    run<3>(tab, bytes, 48, out, p.multiProcessorCount);    // compact x-plane, dense 48-byte records
    run<2>(tab, bytes, 32, out, p.multiProcessorCount);    // one sector
    run<8>(tab, bytes, 128, out, p.multiProcessorCount);   // one full line
    return 0;
}
