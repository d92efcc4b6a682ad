// fp2.cuh — quadratic extension Fp2 = Fp[u]/(u^2 - beta) with beta = -1 (BLS12-381 Fq2: curves/bls12_381/src/fields/fq2.rs:10-24,
// NONRESIDUE = -1), the coordinate field of G2.  Same static interface as Fp<P> so that ec.cuh and the MSM kernels are generic
// over the coordinate field; an element is c0 limbs then c1 limbs (QuadExtField { c0, c1 }, ff/src/fields/models/quadratic_extension.rs:105-113),
// L = 2 * P::L 32-bit words, both halves Montgomery and fully reduced.
//   mul    quadratic_extension.rs:586-614 -> Fp2Config::mul_assign fp2.rs:56-84 (sum_of_products; same value as the Karatsuba form used here)
//   square quadratic_extension.rs:277-325 (complex squaring for beta = -1)
//   inverse quadratic_extension.rs:327-349: (c0 - c1 u) / (c0^2 - beta c1^2)
#pragma once
#include "fp.cuh"

namespace ab200 {

template <class P> struct Fp2 {
    using B = Fp<P>;
    static constexpr int LB = P::L;
    static constexpr int L = 2 * P::L;
    // Base-field multiplication: inlined like everything else.  A called (__noinline__) body would cut the G2 kernels' code size and
    // compile time by 3x, but with it msm_bucket_reduce_kernel<G2> returned wrong sums on the B200 for every bucket index >= 2 (the
    // double-and-add of the chunk offset) while the same source passed on the host and in the element-wise kernels; the inlined
    // build passes the whole matrix (profiles/r02_g2_debug_matrix.log: default vs variants v3 inline / v4 more calls / v5 ptxas -O1).
    // AB_FP2_NOINLINE_MUL re-enables the call form for experiments.
#if defined(__CUDACC__) && defined(AB_FP2_NOINLINE_MUL)
    static __host__ __device__ __noinline__ void bmul(uint32_t *r, const uint32_t *a, const uint32_t *b) { B::mul(r, a, b); }
#else
    static AB_HD void bmul(uint32_t *r, const uint32_t *a, const uint32_t *b) { B::mul(r, a, b); }
#endif
    static AB_HD void add(uint32_t *r, const uint32_t *a, const uint32_t *b) { B::add(r, a, b); B::add(r + LB, a + LB, b + LB); }
    static AB_HD void sub(uint32_t *r, const uint32_t *a, const uint32_t *b) { B::sub(r, a, b); B::sub(r + LB, a + LB, b + LB); }
    static AB_HD void dbl(uint32_t *r, const uint32_t *a) { B::dbl(r, a); B::dbl(r + LB, a + LB); }
    static AB_HD void neg(uint32_t *r, const uint32_t *a) { B::neg(r, a); B::neg(r + LB, a + LB); }
    static AB_HD void cneg(uint32_t *r, const uint32_t *a, bool f) { B::cneg(r, a, f); B::cneg(r + LB, a + LB, f); }
    // (a0 + a1 u)(b0 + b1 u) = (a0 b0 - a1 b1) + ((a0 + a1)(b0 + b1) - a0 b0 - a1 b1) u      [3 base multiplications]
    static AB_HD void mul(uint32_t *r, const uint32_t *a, const uint32_t *b) {
        uint32_t v0[LB], v1[LB], sa[LB], sb[LB];
        bmul(v0, a, b);
        bmul(v1, a + LB, b + LB);
        B::add(sa, a, a + LB);
        B::add(sb, b, b + LB);
        bmul(sa, sa, sb);
        B::sub(sa, sa, v0);
        B::sub(r + LB, sa, v1);
        B::sub(r, v0, v1);
    }
    // (a0 + a1 u)^2 = (a0 + a1)(a0 - a1) + 2 a0 a1 u
    static AB_HD void sqr(uint32_t *r, const uint32_t *a) {
        uint32_t s[LB], d[LB], m[LB];
        B::add(s, a, a + LB);
        B::sub(d, a, a + LB);
        bmul(m, a, a + LB);
        bmul(r, s, d);
        B::dbl(r + LB, m);
    }
    static AB_HD void set_one(uint32_t *r) { B::set_one(r); B::set_zero(r + LB); }
    static AB_HD void set_zero(uint32_t *r) { B::set_zero(r); B::set_zero(r + LB); }
    static AB_HD void inv_lowlat(uint32_t *r, const uint32_t *a) { inv(r, a); }
    static AB_HD void inv(uint32_t *r, const uint32_t *a) {
        uint32_t n[LB], t[LB];
        bmul(n, a, a);
        bmul(t, a + LB, a + LB);
        B::add(n, n, t);       // norm = c0^2 + c1^2
        B::inv(n, n);
        bmul(r, a, n);
        bmul(t, a + LB, n);
        B::neg(r + LB, t);
    }
};

}  // namespace ab200
