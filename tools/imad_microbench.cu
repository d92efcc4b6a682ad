// Integer-multiply pipe microbenchmark for sm_100a: decides the Montgomery formulation
// (mad.lo/.hi carry chains vs mad.wide.u32) and gives the measured IMAD roofline denominator.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_microbench imad_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
    uint32_t x0 = a, x1 = b, x2 = a ^ b, x3 = a + b, x4 = a - b, x5 = ~a, x6 = ~b, x7 = a * 5;
    uint64_t w0 = a, w1 = b, w2 = a ^ b, w3 = a + b;
    double d0 = a * 1e-9, d1 = b * 1e-9, d2 = d0 + 1.0, d3 = d1 + 1.0, da = 1.0 + a * 1e-12, db = b * 1e-12, dc = 4503599627370496.0;
    #pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
        if (MODE == 0) { // mad.lo, 8 independent chains
            asm volatile("mad.lo.u32 %0, %0, %8, %9; mad.lo.u32 %1, %1, %8, %9; mad.lo.u32 %2, %2, %8, %9; mad.lo.u32 %3, %3, %8, %9;"
                         "mad.lo.u32 %4, %4, %8, %9; mad.lo.u32 %5, %5, %8, %9; mad.lo.u32 %6, %6, %8, %9; mad.lo.u32 %7, %7, %8, %9;"
                         : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(a), "r"(b));
        } else if (MODE == 1) { // mad.hi
            asm volatile("mad.hi.u32 %0, %0, %8, %9; mad.hi.u32 %1, %1, %8, %9; mad.hi.u32 %2, %2, %8, %9; mad.hi.u32 %3, %3, %8, %9;"
                         "mad.hi.u32 %4, %4, %8, %9; mad.hi.u32 %5, %5, %8, %9; mad.hi.u32 %6, %6, %8, %9; mad.hi.u32 %7, %7, %8, %9;"
                         : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(a), "r"(b));
        } else if (MODE == 2) { // Montgomery-row pattern: one carry chain of lo/hi pairs over 8 accumulators (b varies per op pair)
            asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %10, %9, %2; madc.hi.cc.u32 %3, %10, %9, %3;"
                         "madc.lo.cc.u32 %4, %11, %9, %4; madc.hi.cc.u32 %5, %11, %9, %5; madc.lo.cc.u32 %6, %12, %9, %6; madc.hi.u32 %7, %12, %9, %7;"
                         : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(a), "r"(b), "r"(a ^ 1), "r"(a ^ 2), "r"(a ^ 3));
        } else if (MODE == 3) { // two independent carry chains interleaved is impossible (one CC) -> two rows back to back
            asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %10, %9, %2; madc.hi.u32 %3, %10, %9, %3;"
                         "mad.lo.cc.u32 %4, %11, %9, %4; madc.hi.cc.u32 %5, %11, %9, %5; madc.lo.cc.u32 %6, %12, %9, %6; madc.hi.u32 %7, %12, %9, %7;"
                         : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(a), "r"(b), "r"(a ^ 1), "r"(a ^ 2), "r"(a ^ 3));
        } else if (MODE == 4) { // mad.wide.u32, 4 chains, multiplicand = low word of a neighbouring chain (not hoistable), 8 per iter
            asm volatile("mad.wide.u32 %0, %4, %5, %0; mad.wide.u32 %1, %6, %5, %1; mad.wide.u32 %2, %7, %5, %2; mad.wide.u32 %3, %8, %5, %3;"
                         : "+l"(w0), "+l"(w1), "+l"(w2), "+l"(w3) : "r"((uint32_t)w1), "r"(b), "r"((uint32_t)w2), "r"((uint32_t)w3), "r"((uint32_t)w0));
            asm volatile("mad.wide.u32 %0, %4, %5, %0; mad.wide.u32 %1, %6, %5, %1; mad.wide.u32 %2, %7, %5, %2; mad.wide.u32 %3, %8, %5, %3;"
                         : "+l"(w0), "+l"(w1), "+l"(w2), "+l"(w3) : "r"((uint32_t)w1), "r"(a), "r"((uint32_t)w2), "r"((uint32_t)w3), "r"((uint32_t)w0));
        } else if (MODE == 5) { // mad.lo with immediate multiplicand
            asm volatile("mad.lo.u32 %0, %0, 0x9e3779b9, %8; mad.lo.u32 %1, %1, 0x9e3779b9, %8; mad.lo.u32 %2, %2, 0x9e3779b9, %8; mad.lo.u32 %3, %3, 0x9e3779b9, %8;"
                         "mad.lo.u32 %4, %4, 0x85ebca6b, %8; mad.lo.u32 %5, %5, 0x85ebca6b, %8; mad.lo.u32 %6, %6, 0x85ebca6b, %8; mad.lo.u32 %7, %7, 0x85ebca6b, %8;"
                         : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(a));
        } else if (MODE == 6) { // mad.hi.cc with immediate multiplicand (the k*p[j] half of a Montgomery row)
            asm volatile("mad.lo.cc.u32 %0, %8, 0x9e3779b9, %0; madc.hi.cc.u32 %1, %8, 0x9e3779b9, %1; madc.lo.cc.u32 %2, %8, 0x85ebca6b, %2; madc.hi.cc.u32 %3, %8, 0x85ebca6b, %3;"
                         "madc.lo.cc.u32 %4, %8, 0xc2b2ae35, %4; madc.hi.cc.u32 %5, %8, 0xc2b2ae35, %5; madc.lo.cc.u32 %6, %8, 0x27d4eb2f, %6; madc.hi.u32 %7, %8, 0x27d4eb2f, %7;"
                         : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(a));
        } else if (MODE == 7) { // 8 mad.lo + 8 independent add (ALU pipe) : co-issue test
            asm volatile("mad.lo.u32 %0, %0, %8, %9; add.u32 %4, %4, %8; mad.lo.u32 %1, %1, %8, %9; add.u32 %5, %5, %9; mad.lo.u32 %2, %2, %8, %9; add.u32 %6, %6, %8; mad.lo.u32 %3, %3, %8, %9; add.u32 %7, %7, %9;"
                         "mad.lo.u32 %0, %0, %9, %8; add.u32 %4, %4, %9; mad.lo.u32 %1, %1, %9, %8; add.u32 %5, %5, %8; mad.lo.u32 %2, %2, %9, %8; add.u32 %6, %6, %9; mad.lo.u32 %3, %3, %9, %8; add.u32 %7, %7, %8;"
                         : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(a), "r"(b));
        } else if (MODE == 8) { // 4 x (mad.lo.cc + madc.hi.cc) pairs with loop-variant multiplicand: the Montgomery row as ptxas sees it
            asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %10, %9, %2; madc.hi.cc.u32 %3, %10, %9, %3;"
                         "madc.lo.cc.u32 %4, %11, %9, %4; madc.hi.cc.u32 %5, %11, %9, %5; madc.lo.cc.u32 %6, %12, %9, %6; madc.hi.u32 %7, %12, %9, %7;"
                         : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(x7 | 1), "r"(b), "r"(x1 | 1), "r"(x3 | 1), "r"(x5 | 1));
        } else if (MODE == 9) { // mul.wide (no accumulate), loop-variant
            asm volatile("mul.wide.u32 %0, %4, %5; mul.wide.u32 %1, %6, %5; mul.wide.u32 %2, %7, %5; mul.wide.u32 %3, %8, %5;" : "=l"(w0), "=l"(w1), "=l"(w2), "=l"(w3) : "r"((uint32_t)w1 | 1), "r"(b), "r"((uint32_t)w2 | 1), "r"((uint32_t)w3 | 1), "r"((uint32_t)w0 | 1));
            asm volatile("mul.wide.u32 %0, %4, %5; mul.wide.u32 %1, %6, %5; mul.wide.u32 %2, %7, %5; mul.wide.u32 %3, %8, %5;" : "=l"(w0), "=l"(w1), "=l"(w2), "=l"(w3) : "r"((uint32_t)w1 | 1), "r"(a), "r"((uint32_t)w2 | 1), "r"((uint32_t)w3 | 1), "r"((uint32_t)w0 | 1));
        } else if (MODE == 10) { // DFMA x8 independent chains
            asm volatile("fma.rn.f64 %0, %0, %4, %5; fma.rn.f64 %1, %1, %4, %5; fma.rn.f64 %2, %2, %4, %5; fma.rn.f64 %3, %3, %4, %5;"
                         "fma.rn.f64 %0, %0, %5, %4; fma.rn.f64 %1, %1, %5, %4; fma.rn.f64 %2, %2, %5, %4; fma.rn.f64 %3, %3, %5, %4;"
                         : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3) : "d"(da), "d"(db));
        } else if (MODE == 11) { // DFMA x8 + 8 IMAD.WIDE-pairs interleaved: do the two pipes overlap?
            asm volatile("fma.rn.f64 %0, %0, %4, %5; fma.rn.f64 %1, %1, %4, %5; fma.rn.f64 %2, %2, %4, %5; fma.rn.f64 %3, %3, %4, %5;"
                         "fma.rn.f64 %0, %0, %5, %4; fma.rn.f64 %1, %1, %5, %4; fma.rn.f64 %2, %2, %5, %4; fma.rn.f64 %3, %3, %5, %4;"
                         : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3) : "d"(da), "d"(db));
            asm volatile("mad.wide.u32 %0, %4, %5, %0; mad.wide.u32 %1, %6, %5, %1; mad.wide.u32 %2, %7, %5, %2; mad.wide.u32 %3, %8, %5, %3;"
                         : "+l"(w0), "+l"(w1), "+l"(w2), "+l"(w3) : "r"((uint32_t)w1), "r"(b), "r"((uint32_t)w2), "r"((uint32_t)w3), "r"((uint32_t)w0));
        } else if (MODE == 12) { // FP64 2-instruction 52x52 product (hi via fma.rz, lo via fma): the Emmart split, 4 products per iter
            double h0, h1, h2, h3;
            asm volatile("fma.rz.f64 %0, %4, %5, %6; fma.rz.f64 %1, %4, %7, %6; fma.rz.f64 %2, %4, %8, %6; fma.rz.f64 %3, %4, %9, %6;"
                         : "=d"(h0), "=d"(h1), "=d"(h2), "=d"(h3) : "d"(da), "d"(d0), "d"(dc), "d"(d1), "d"(d2), "d"(d3));
            asm volatile("fma.rz.f64 %0, %4, %0, %5; fma.rz.f64 %1, %4, %1, %6; fma.rz.f64 %2, %4, %2, %7; fma.rz.f64 %3, %4, %3, %8;"
                         : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3) : "d"(da), "d"(h0), "d"(h1), "d"(h2), "d"(h3));
        }
    }
    x0 ^= (uint32_t)__double_as_longlong(d0 + d1 + d2 + d3);
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7 ^ (uint32_t)(w0 ^ w1 ^ w2 ^ w3) ^ (uint32_t)((w0 ^ w1 ^ w2 ^ w3) >> 32);
}

template <int MODE>
int run(const char *name, int mults_per_iter, uint32_t *out, int sms, int threads_per_sm) {
    int blocks = sms * (threads_per_sm / 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, 12345); CHK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(out, 12345 + r); cudaEventRecord(e1); CHK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double ops = (double)blocks * 256 * ITERS * mults_per_iter;
    printf("{\"bench\":\"%s\",\"threads_per_sm\":%d,\"ms\":%.4f,\"Gmul_per_s\":%.1f,\"mul_per_clk_per_sm_at_1965MHz\":%.2f}\n", name, threads_per_sm, best, ops / best / 1e6, ops / (best * 1e-3) / sms / 1.965e9);
    return 0;
}

int main() {
    cudaDeviceProp p; CHK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n", p.name, sms, p.clockRate);
    uint32_t *out; CHK(cudaMalloc(&out, (size_t)sms * 2048 * 4));
    for (int tps : {256, 512, 1024, 2048}) {
        run<0>("mad.lo.u32 x8 indep", 8, out, sms, tps);
        run<1>("mad.hi.u32 x8 indep", 8, out, sms, tps);
        run<2>("mad.lo.cc/madc.hi.cc chain of 8", 8, out, sms, tps);
        run<3>("two cc chains of 4", 8, out, sms, tps);
        run<4>("mad.wide.u32 x8 variant ops", 8, out, sms, tps);
        run<5>("mad.lo.u32 imm x8", 8, out, sms, tps);
        run<6>("mad.lo/hi.cc imm chain of 8", 8, out, sms, tps);
        run<7>("16 mad.lo + 16 add.u32 interleaved", 16, out, sms, tps);
        run<8>("4 wide products as lo.cc/hi.cc pairs, variant ops", 4, out, sms, tps);
        run<9>("mul.wide.u32 x8 variant", 8, out, sms, tps);
        run<10>("DFMA x8", 8, out, sms, tps);
        run<11>("DFMA x8 + mad.wide x4 (count DFMA)", 8, out, sms, tps);
        run<12>("DFMA pairs rz (8 DFMA)", 8, out, sms, tps);
    }
    return 0;
}
