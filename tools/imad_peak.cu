// IMAD.WIDE issue-rate microbenchmark with a SASS-verifiable instruction count (settles the denominator of every IMAD
// roofline fraction in bench.py).  Each mode's loop body is straight PTX whose SASS opcode count is checked at build time
// (tools/imad_peak_sass.py counts IMAD.WIDE per loop iteration from `cuobjdump -sass`), so ops/iteration is not a guess.
//   mode 0: 16 independent  mad.wide.u32 acc, lo(acc), b, acc      (IMAD.WIDE.U32 with 64-bit addend, no carry)
//   mode 1: 2 x carry chains of 4 (mad.lo.cc / madc.hi.cc pairs -> IMAD.WIDE.U32 / IMAD.WIDE.U32.X): the Montgomery row form
//   mode 2: 16 independent  mul.wide.u32 with the result folded by xor into its own multiplicand (no addend)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_peak imad_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 8192
template <int MODE> __global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed) {
    uint64_t w[16];
    uint32_t b = seed * 2654435761u + blockIdx.x, x[8];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = ((uint64_t)(seed + i) << 32) | (threadIdx.x * 2654435761u + i);
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i * seed;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((uint32_t)w[i]), "r"(b));
        } else if (MODE == 1) {
            asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %10, %9, %2; madc.hi.cc.u32 %3, %10, %9, %3;"
                         "madc.lo.cc.u32 %4, %11, %9, %4; madc.hi.cc.u32 %5, %11, %9, %5; madc.lo.cc.u32 %6, %12, %9, %6; madc.hi.u32 %7, %12, %9, %7;"
                         : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7])
                         : "r"((uint32_t)w[0]), "r"(b), "r"((uint32_t)w[1]), "r"((uint32_t)w[2]), "r"((uint32_t)w[3]));
            asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %10, %9, %2; madc.hi.cc.u32 %3, %10, %9, %3;"
                         "madc.lo.cc.u32 %4, %11, %9, %4; madc.hi.cc.u32 %5, %11, %9, %5; madc.lo.cc.u32 %6, %12, %9, %6; madc.hi.u32 %7, %12, %9, %7;"
                         : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7])
                         : "r"((uint32_t)w[4]), "r"(b ^ 0x5555u), "r"((uint32_t)w[5]), "r"((uint32_t)w[6]), "r"((uint32_t)w[7]));
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                uint64_t t;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"((uint32_t)w[i]), "r"(b));
                w[i] = t;
            }
        }
    }
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) acc ^= w[i];
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)acc ^ (uint32_t)(acc >> 32);
}
template <int MODE> static void run(const char *name, int wide_per_iter, uint32_t *out, int sms, int tps, double ghz) {
    int blocks = sms * (tps / 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, 1); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(out, 7 + r); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double ops = (double)blocks * 256 * ITERS * wide_per_iter;
    printf("{\"bench\":\"%s\",\"threads_per_sm\":%d,\"wide_per_iter_sass\":%d,\"ms\":%.4f,\"T_wide_per_s\":%.3f,\"wide_per_clk_per_sm\":%.2f}\n", name, tps,
           wide_per_iter, best, ops / best / 1e9, ops / (best * 1e-3) / sms / (ghz * 1e9));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; double ghz = p.clockRate / 1e6;
    printf("{\"device\":\"%s\",\"sms\":%d,\"clock_ghz\":%.3f}\n", p.name, sms, ghz);
    uint32_t *out; cudaMalloc(&out, (size_t)sms * 2048 * 4);
    for (int tps : {512, 1024, 2048}) {
        run<0>("mad.wide.u32 x16 independent (IMAD.WIDE.U32, 64-bit addend)", 16, out, sms, tps, ghz);
        run<1>("2 carry chains of 4 lo/hi pairs (IMAD.WIDE.U32[.X])", 8, out, sms, tps, ghz);
        run<2>("mul.wide.u32 x16 independent (IMAD.WIDE.U32, RZ addend)", 16, out, sms, tps, ghz);
    }
    return 0;
}
