// fp.cuh — Montgomery prime-field arithmetic for sm_100a, 32-bit limbs, L in {8, 12}.
//
// Device counterpart of ark_ff's `Fp<MontBackend<T,N>,N>` (N = L/2 u64 limbs):
//   mul        ff/src/fields/models/fp/montgomery_backend.rs:179-246  (CIOS, one final conditional subtract)
//   add/sub    :128-148      dbl :151-161      neg :164-171
//   into_bigint :392-412     from_bigint :380-390
// Values are always fully reduced (< p) on entry and exit, exactly like the reference, so every
// intermediate is bit-identical to what ark_ff holds (in-memory layout: little-endian u64 limbs ==
// little-endian u32 limb pairs).
//
// Multiplication: word-serial Montgomery with the odd/even column split — products a[2k]*b_i land on
// word pairs (2k,2k+1), products a[2k+1]*b_i on pairs (2k+1,2k+2); keeping the two families in two
// accumulators lets each family be ONE carry chain of (mad.lo.cc, madc.hi.cc) pairs, which ptxas fuses
// into a single IMAD.WIDE.U32(.X) per 32x32 product on sm_100a (checked with cuobjdump).  2L^2 + L
// wide MADs per modmul: 136 (L=8), 300 (L=12) — the figure SURVEY.md §8(d) uses.
//
// The PTX primitives have a host emulation (carry flag in a thread_local) so the very same algorithm
// text can be exercised on the CPU build box by tools/host_selftest.cu; device code never takes that path.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define AB_HD __host__ __device__ __forceinline__
#define AB_D __device__ __forceinline__
#else
#define AB_HD inline
#define AB_D inline
#endif

namespace ab200 {

// ------------------------------------------------------------------------------------------------
// PTX carry-chain primitives
// ------------------------------------------------------------------------------------------------
namespace ptx {
#ifndef __CUDA_ARCH__
static thread_local uint32_t host_cc = 0;
#endif

AB_HD uint32_t add_cc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    uint64_t t = (uint64_t)a + b; host_cc = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
AB_HD uint32_t addc_cc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    uint64_t t = (uint64_t)a + b + host_cc; host_cc = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
AB_HD uint32_t addc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return a + b + host_cc;
#endif
}
AB_HD uint32_t sub_cc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    uint64_t t = (uint64_t)a - b; host_cc = (uint32_t)(t >> 32) & 1; return (uint32_t)t;   // cc = borrow
#endif
}
AB_HD uint32_t subc_cc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    uint64_t t = (uint64_t)a - b - host_cc; host_cc = (uint32_t)(t >> 32) & 1; return (uint32_t)t;
#endif
}
AB_HD uint32_t subc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return a - b - host_cc;
#endif
}
AB_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
AB_HD uint32_t mul_hi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
AB_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    uint64_t t = (uint64_t)(uint32_t)(a * b) + c; host_cc = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
AB_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    uint64_t t = (uint64_t)(uint32_t)(a * b) + c + host_cc; host_cc = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
AB_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    uint64_t t = (((uint64_t)a * b) >> 32) + c + host_cc; host_cc = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
AB_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    return (uint32_t)((((uint64_t)a * b) >> 32) + c + host_cc);
#endif
}
}  // namespace ptx

// ------------------------------------------------------------------------------------------------
// Field parameter packs.  Constants are constexpr-function tables so that after full unrolling every
// modulus limb is an immediate operand of the IMAD (no constant-bank or register traffic).
// Moduli: curves/bls12_381/src/fields/{fq,fr}.rs, curves/bn254/src/fields/{fq,fr}.rs.
// R (= ONE), R2 and INV follow montgomery_backend.rs:21-24,520-538; they are re-derived and asserted by
// tools/gen_consts.py and by tests/test_host_selftest.py.
// ------------------------------------------------------------------------------------------------
#define AB_TABLE(NAME, ...)                                                                  \
    static constexpr AB_HD uint32_t NAME(int i) {                                            \
        const uint32_t t[] = {__VA_ARGS__};                                                  \
        return t[i];                                                                         \
    }

#include "field_consts.inc"

// Same fields, but Fp::mul keeps its row pairs in a real loop instead of unrolling all L rows: ~3x less code per multiplication,
// so a 10-multiplication point addition fits the instruction cache (the fully unrolled madd body is ~53 KB of SASS and ncu shows
// `no_instruction` stalls on it).  Used by the MSM accumulation kernel only.
struct BlsFqRolled : BlsFq { static constexpr bool ROLLED = true; };
struct BnFqRolled : BnFq { static constexpr bool ROLLED = true; };

// ------------------------------------------------------------------------------------------------
// Generic helpers on raw limb arrays
// ------------------------------------------------------------------------------------------------
template <int L> AB_HD void limbs_copy(uint32_t *r, const uint32_t *a) {
#pragma unroll
    for (int i = 0; i < L; i++) r[i] = a[i];
}
template <int L> AB_HD bool limbs_is_zero(const uint32_t *a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < L; i++) o |= a[i];
    return o == 0;
}
template <int L> AB_HD bool limbs_eq(const uint32_t *a, const uint32_t *b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < L; i++) o |= a[i] ^ b[i];
    return o == 0;
}

// The trivial multipliers 1 and 2^32-1 of "Montgomery-friendly" moduli (BLS12-381 Fr = ...ffffffff00000001, INV = -1) are read
// from constant memory instead of being immediates, so that ptxas keeps every (mad.lo.cc, madc.hi.cc) pair a fused
// IMAD.WIDE.U32(.X): with strength-reduced forms (m = -T[0], m*1 -> add, m*(2^32-1) -> shift/sub) the carry chains start with
// IADD3s and the remaining pairs of the row come out as IMAD.X + IMAD.HI.U32.X — two issues on the multiplier pipe instead of
// one.  Round 1 shipped that form: 66 fused + 48 unfused pairs per Fr multiplication; now 123 fused (cuobjdump counts in
// profiles/r02_sass_opcode_counts.txt).
#ifdef __CUDACC__
static __constant__ uint32_t ab_opaque_words[2] = {1u, 0xffffffffu};
#endif

template <class P> struct Fp {
    static constexpr int L = P::L;
    // -p^-1 mod 2^32.  For p = 1 (mod 2^32) this is 2^32-1 and m = -T[0]; written as a negation, ptxas stops fusing ALL m*p[j]
    // products of the row (IMAD.X + IMAD.HI.U32.X pairs instead of IMAD.WIDE.U32.X: 48 of 115 per multiplication for BLS12-381
    // Fr), so the multiplier stays an opaque constant-memory word and m is a real product.
    static AB_HD uint32_t inv_word() {
#ifdef __CUDA_ARCH__
        if (P::INV32 == 0xffffffffu) return ab_opaque_words[1];
#endif
        return P::INV32;
    }
    // modulus word I as a multiplier operand
    template <int I> static AB_HD uint32_t mod_word() {
#ifdef __CUDA_ARCH__
        if (P::MOD(I) == 1u) return ab_opaque_words[0];
        if (P::MOD(I) == 0xffffffffu) return ab_opaque_words[1];
#endif
        return P::MOD(I);
    }

    // r = (a >= p) ? a - p : a        (`subtract_modulus`, ff/src/fields/models/fp/mod.rs:140-155)
    static AB_HD void reduce_once(uint32_t *a) {
        uint32_t t[L];
        t[0] = ptx::sub_cc(a[0], P::MOD(0));
#pragma unroll
        for (int i = 1; i < L; i++) t[i] = ptx::subc_cc(a[i], P::MOD(i));
        uint32_t borrow = ptx::subc(0u, 0u);  // 0xffffffff if a < p
#pragma unroll
        for (int i = 0; i < L; i++) a[i] = borrow ? a[i] : t[i];
    }

    // montgomery_backend.rs:128-138 — the modulus has a spare bit, so a+b never carries out of L words.
    static AB_HD void add(uint32_t *r, const uint32_t *a, const uint32_t *b) {
        uint32_t s[L];
        s[0] = ptx::add_cc(a[0], b[0]);
#pragma unroll
        for (int i = 1; i < L - 1; i++) s[i] = ptx::addc_cc(a[i], b[i]);
        s[L - 1] = ptx::addc(a[L - 1], b[L - 1]);
        reduce_once(s);
        limbs_copy<L>(r, s);
    }
    static AB_HD void dbl(uint32_t *r, const uint32_t *a) { add(r, a, a); }  // :151-161

    // montgomery_backend.rs:141-148: if b > a add p first, then subtract.
    static AB_HD void sub(uint32_t *r, const uint32_t *a, const uint32_t *b) {
        uint32_t d[L];
        d[0] = ptx::sub_cc(a[0], b[0]);
#pragma unroll
        for (int i = 1; i < L; i++) d[i] = ptx::subc_cc(a[i], b[i]);
        uint32_t borrow = ptx::subc(0u, 0u);  // all-ones mask if a < b
        r[0] = ptx::add_cc(d[0], P::MOD(0) & borrow);
#pragma unroll
        for (int i = 1; i < L - 1; i++) r[i] = ptx::addc_cc(d[i], P::MOD(i) & borrow);
        r[L - 1] = ptx::addc(d[L - 1], P::MOD(L - 1) & borrow);
    }
    // :164-171
    static AB_HD void neg(uint32_t *r, const uint32_t *a) {
        uint32_t nz = 0;
#pragma unroll
        for (int i = 0; i < L; i++) nz |= a[i];
        uint32_t mask = nz ? 0xffffffffu : 0u;
        uint32_t t[L];
        t[0] = ptx::sub_cc(P::MOD(0), a[0]);
#pragma unroll
        for (int i = 1; i < L - 1; i++) t[i] = ptx::subc_cc(P::MOD(i), a[i]);
        t[L - 1] = ptx::subc(P::MOD(L - 1), a[L - 1]);
#pragma unroll
        for (int i = 0; i < L; i++) r[i] = t[i] & mask;
    }
    // conditional negate: neg_flag != 0 -> -a
    static AB_HD void cneg(uint32_t *r, const uint32_t *a, bool neg_flag) {
        uint32_t t[L];
        neg(t, a);
#pragma unroll
        for (int i = 0; i < L; i++) r[i] = neg_flag ? t[i] : a[i];
    }

    // One row: T += a*bi ; m = T[0]*INV ; T += m*p ; T >>= 32, on the split accumulators (see header).
    // `ev` holds offset-0 words, `od` offset-1 words; on return the roles of the two arrays are swapped.
    template <bool FIRST> static AB_HD void mont_row(uint32_t *ev, uint32_t *od, const uint32_t *a, uint32_t bi) {
        if (FIRST) {
#pragma unroll
            for (int j = 0; j < L; j += 2) {
                od[j] = ptx::mul_lo(a[j + 1], bi);
                od[j + 1] = ptx::mul_hi(a[j + 1], bi);
                ev[j] = ptx::mul_lo(a[j], bi);
                ev[j + 1] = ptx::mul_hi(a[j], bi);
            }
        } else {
            // previous row left: T = (od_prev>>32) + ev_prev with od_prev[0] == 0; our `ev` is the old odd array
            // (already offset 0) and `od` the old even array, which must drop two words to become offset 1.
            ev[0] = ptx::add_cc(ev[0], od[1]);
#pragma unroll
            for (int j = 0; j < L - 2; j += 2) {
                od[j] = ptx::madc_lo_cc(a[j + 1], bi, od[j + 2]);
                od[j + 1] = ptx::madc_hi_cc(a[j + 1], bi, od[j + 3]);
            }
            od[L - 2] = ptx::madc_lo_cc(a[L - 1], bi, 0u);
            od[L - 1] = ptx::madc_hi(a[L - 1], bi, 0u);
            ev[0] = ptx::mad_lo_cc(a[0], bi, ev[0]);
            ev[1] = ptx::madc_hi_cc(a[0], bi, ev[1]);
#pragma unroll
            for (int j = 2; j < L; j += 2) {
                ev[j] = ptx::madc_lo_cc(a[j], bi, ev[j]);
                ev[j + 1] = ptx::madc_hi_cc(a[j], bi, ev[j + 1]);
            }
            od[L - 1] = ptx::addc(od[L - 1], 0u);
        }
        // m = T[0] * (-p^-1) mod 2^32; for moduli with p = 1 (mod 2^32) (BLS12-381 Fr) that is simply -T[0]
        const uint32_t m = ev[0] * inv_word();
        od[0] = ptx::mad_lo_cc(mod_word<1>(), m, od[0]);
        od[1] = ptx::madc_hi_cc(mod_word<1>(), m, od[1]);
#pragma unroll
        for (int j = 2; j < L; j += 2) {
            od[j] = ptx::madc_lo_cc(P::MOD(j + 1), m, od[j]);
            od[j + 1] = ptx::madc_hi_cc(P::MOD(j + 1), m, od[j + 1]);
        }
        ev[0] = ptx::mad_lo_cc(mod_word<0>(), m, ev[0]);
        ev[1] = ptx::madc_hi_cc(mod_word<0>(), m, ev[1]);
#pragma unroll
        for (int j = 2; j < L; j += 2) {
            ev[j] = ptx::madc_lo_cc(P::MOD(j), m, ev[j]);
            ev[j + 1] = ptx::madc_hi_cc(P::MOD(j), m, ev[j + 1]);
        }
        od[L - 1] = ptx::addc(od[L - 1], 0u);
    }

    // r = a*b*R^-1 mod p, fully reduced.  r may alias a or b.
    static AB_HD void mul(uint32_t *r, const uint32_t *a, const uint32_t *b) {
        uint32_t ev[L], od[L];
        mont_row<true>(ev, od, a, b[0]);
        mont_row<false>(od, ev, a, b[1]);
        if (P::ROLLED) {
            uint32_t bb[L];  // dynamically indexed below -> lives in (L1-resident) local memory
            limbs_copy<L>(bb, b);
#pragma unroll 1
            for (int i = 2; i < L; i += 2) {
                mont_row<false>(ev, od, a, bb[i]);
                mont_row<false>(od, ev, a, bb[i + 1]);
            }
        } else {
#pragma unroll
            for (int i = 2; i < L; i += 2) {
                mont_row<false>(ev, od, a, b[i]);
                mont_row<false>(od, ev, a, b[i + 1]);
            }
        }
        // after an even number of rows: T = (od>>32) + ev
        uint32_t t[L];
        t[0] = ptx::add_cc(ev[0], od[1]);
#pragma unroll
        for (int i = 1; i < L - 1; i++) t[i] = ptx::addc_cc(ev[i], od[i + 1]);
        t[L - 1] = ptx::addc(ev[L - 1], 0u);
        reduce_once(t);
        limbs_copy<L>(r, t);
    }
    // square_in_place (montgomery_backend.rs:248-317).  Measured on B200: the symmetric squaring below (sqr_sos, 234 instead
    // of 300 wide MADs for L = 12) made msm_accumulate_kernel 2 % SLOWER (338 -> 345 ms @2^26: three 2L-word temporaries
    // raise register pressure and the extra carry-ripple adds cost issue slots), so sqr stays mul(a, a) — which is also what
    // the reference's asm path does (ff-asm/src/lib.rs:132).  sqr_sos is kept, tested, for latency-bound single-thread code.
    static AB_HD void sqr(uint32_t *r, const uint32_t *a) { mul(r, a, a); }

    // Symmetric squaring: t = a^2 by symmetry (L(L-1)/2 off-diagonal products, doubled, + L diagonal squares) followed by
    // L word-serial reduction rows.  Same value as mul(a, a).
    static AB_HD void sqr_sos(uint32_t *r, const uint32_t *a) {
        // ---- off-diagonal sum S = sum_{i<j} a_i a_j 2^(32(i+j)), split by the parity of i+j into two carry-chain grids
        uint32_t E[2 * L], O[2 * L];  // E: pairs (2k, 2k+1); O: pairs (2k+1, 2k+2)
#pragma unroll
        for (int i = 0; i < 2 * L; i++) { E[i] = 0u; O[i] = 0u; }
#pragma unroll
        for (int i = 0; i < L - 1; i++) {
            // i + j odd  -> O grid: j = i+1, i+3, ...
            {
                int w = 0;
                bool first = true;
#pragma unroll
                for (int j = i + 1; j < L; j += 2) {
                    w = i + j;
                    O[w] = first ? ptx::mad_lo_cc(a[i], a[j], O[w]) : ptx::madc_lo_cc(a[i], a[j], O[w]);
                    O[w + 1] = ptx::madc_hi_cc(a[i], a[j], O[w + 1]);
                    first = false;
                }
                if (w + 2 < 2 * L) O[w + 2] = ptx::addc(O[w + 2], 0u);  // the word above this row's chain is still zero
            }
            // i + j even -> E grid: j = i+2, i+4, ...
            if (i + 2 < L) {
                int w = 0;
                bool first = true;
#pragma unroll
                for (int j = i + 2; j < L; j += 2) {
                    w = i + j;
                    E[w] = first ? ptx::mad_lo_cc(a[i], a[j], E[w]) : ptx::madc_lo_cc(a[i], a[j], E[w]);
                    E[w + 1] = ptx::madc_hi_cc(a[i], a[j], E[w + 1]);
                    first = false;
                }
                if (w + 2 < 2 * L) E[w + 2] = ptx::addc(E[w + 2], 0u);
            }
        }
        // ---- t = 2*(E + O) + sum_i a_i^2 2^(64 i)
        uint32_t t[2 * L];
        t[0] = ptx::add_cc(E[0], O[0]);
#pragma unroll
        for (int i = 1; i < 2 * L - 1; i++) t[i] = ptx::addc_cc(E[i], O[i]);
        t[2 * L - 1] = ptx::addc(E[2 * L - 1], O[2 * L - 1]);
        t[0] = ptx::add_cc(t[0], t[0]);
#pragma unroll
        for (int i = 1; i < 2 * L - 1; i++) t[i] = ptx::addc_cc(t[i], t[i]);
        t[2 * L - 1] = ptx::addc(t[2 * L - 1], t[2 * L - 1]);
        t[0] = ptx::mad_lo_cc(a[0], a[0], t[0]);
        t[1] = ptx::madc_hi_cc(a[0], a[0], t[1]);
#pragma unroll
        for (int i = 1; i < L; i++) {
            t[2 * i] = ptx::madc_lo_cc(a[i], a[i], t[2 * i]);
            t[2 * i + 1] = (i == L - 1) ? ptx::madc_hi(a[i], a[i], t[2 * i + 1]) : ptx::madc_hi_cc(a[i], a[i], t[2 * i + 1]);
        }
        redc_wide(r, t);
    }

    // One reduction row on the split accumulators (cf. mont_row): T = (T + m*p) >> 32 with m = T[0]*INV, then the next
    // high word of the 2L-word input enters at the top.  FIRST: ev = low words, od = 0, nothing to shift in yet.
    template <bool FIRST> static AB_HD void redc_row(uint32_t *ev, uint32_t *od, uint32_t hi_word) {
        if (!FIRST) {
            ev[0] = ptx::add_cc(ev[0], od[1]);
#pragma unroll
            for (int j = 0; j < L - 2; j++) od[j] = ptx::addc_cc(od[j + 2], 0u);
            od[L - 2] = ptx::addc_cc(hi_word, 0u);
            od[L - 1] = ptx::addc(0u, 0u);
        }
        const uint32_t m = ev[0] * inv_word();
        od[0] = ptx::mad_lo_cc(mod_word<1>(), m, od[0]);
        od[1] = ptx::madc_hi_cc(mod_word<1>(), m, od[1]);
#pragma unroll
        for (int j = 2; j < L; j += 2) {
            od[j] = ptx::madc_lo_cc(P::MOD(j + 1), m, od[j]);
            od[j + 1] = ptx::madc_hi_cc(P::MOD(j + 1), m, od[j + 1]);
        }
        ev[0] = ptx::mad_lo_cc(mod_word<0>(), m, ev[0]);
        ev[1] = ptx::madc_hi_cc(mod_word<0>(), m, ev[1]);
#pragma unroll
        for (int j = 2; j < L; j += 2) {
            ev[j] = ptx::madc_lo_cc(P::MOD(j), m, ev[j]);
            ev[j + 1] = ptx::madc_hi_cc(P::MOD(j), m, ev[j + 1]);
        }
        od[L - 1] = ptx::addc(od[L - 1], 0u);
    }

    // r = t * R^-1 mod p for a 2L-word t < p*R (Montgomery reduction, separated from the multiplication)
    static AB_HD void redc_wide(uint32_t *r, const uint32_t *t) {
        uint32_t ev[L], od[L];
#pragma unroll
        for (int i = 0; i < L; i++) { ev[i] = t[i]; od[i] = 0u; }
        redc_row<true>(ev, od, 0u);
        redc_row<false>(od, ev, t[L]);
#pragma unroll
        for (int i = 2; i < L; i += 2) {
            redc_row<false>(ev, od, t[L + i - 1]);
            redc_row<false>(od, ev, t[L + i]);
        }
        // after an even number of rows: T = (od >> 32) + ev, plus the last high word at the top
        uint32_t u[L];
        u[0] = ptx::add_cc(ev[0], od[1]);
#pragma unroll
        for (int i = 1; i < L - 1; i++) u[i] = ptx::addc_cc(ev[i], od[i + 1]);
        u[L - 1] = ptx::addc(ev[L - 1], t[2 * L - 1]);
        reduce_once(u);
        limbs_copy<L>(r, u);
    }

    // Montgomery -> canonical: multiply by the integer 1 (montgomery_backend.rs:392-412: N REDC rounds).
    static AB_HD void from_mont(uint32_t *r, const uint32_t *a) {
        uint32_t one[L];
#pragma unroll
        for (int i = 0; i < L; i++) one[i] = (i == 0) ? 1u : 0u;
        mul(r, a, one);
    }
    // canonical -> Montgomery: multiply by R^2 (:380-390)
    static AB_HD void to_mont(uint32_t *r, const uint32_t *a) {
        uint32_t r2[L];
#pragma unroll
        for (int i = 0; i < L; i++) r2[i] = P::R2(i);
        mul(r, a, r2);
    }
    static AB_HD void set_one(uint32_t *r) {
#pragma unroll
        for (int i = 0; i < L; i++) r[i] = P::ONE(i);
    }
    static AB_HD void set_zero(uint32_t *r) {
#pragma unroll
        for (int i = 0; i < L; i++) r[i] = 0u;
    }
    // a^e, e given as `nw` little-endian u32 words (square-and-multiply, MSB first)
    static AB_HD void pow(uint32_t *r, const uint32_t *a, const uint32_t *e, int nw) {
        uint32_t acc[L], base[L];
        set_one(acc);
        limbs_copy<L>(base, a);
        for (int i = nw * 32 - 1; i >= 0; i--) {
            sqr(acc, acc);
            if ((e[i >> 5] >> (i & 31)) & 1) mul(acc, acc, base);
        }
        limbs_copy<L>(r, acc);
    }
    static AB_HD void pow_u64(uint32_t *r, const uint32_t *a, uint64_t e) {
        uint32_t w[2] = {(uint32_t)e, (uint32_t)(e >> 32)};
        pow(r, a, w, 2);
    }
    // a^(p-2)  (value-identical to the reference's binary-Euclid inverse, montgomery_backend.rs:319-378)
    static AB_HD void inv(uint32_t *r, const uint32_t *a) {
        uint32_t e[L];
        e[0] = ptx::sub_cc(P::MOD(0), 2u);
#pragma unroll
        for (int i = 1; i < L; i++) e[i] = ptx::subc_cc(P::MOD(i), 0u);
        pow(r, a, e, L);
    }
    // a^(p-2) for LATENCY-bound callers (one warp inverting for a whole block, msm_k_pair.cuh::block_inverse): 4-bit fixed windows
    // (one multiplication per 4 squarings instead of ~2: 381 + 95 + 14 instead of 381 + 190 for the 381-bit modulus) and the symmetric
    // squaring sqr_sos (78 products in independent chains + reduction: shorter dependency chain than mul(a, a) for a lone warp; it was
    // measured slower only where the multiplier pipe is saturated).  Same value as inv().
    static AB_HD void inv_lowlat(uint32_t *r, const uint32_t *a) {
        uint32_t e[L];
        e[0] = ptx::sub_cc(P::MOD(0), 2u);
#pragma unroll
        for (int i = 1; i < L; i++) e[i] = ptx::subc_cc(P::MOD(i), 0u);
        uint32_t tab[16][L];   // tab[i] = a^i (dynamically indexed: lives in local memory, L1-resident)
        set_one(tab[0]);
        limbs_copy<L>(tab[1], a);
        for (int i = 2; i < 16; i++) mul(tab[i], tab[i - 1], a);
        uint32_t acc[L];
        set_one(acc);
        bool started = false;
        for (int nib = 8 * L - 1; nib >= 0; nib--) {
            const uint32_t d = (e[nib >> 3] >> ((nib & 7) * 4)) & 15u;
            if (started) {
                sqr_sos(acc, acc); sqr_sos(acc, acc); sqr_sos(acc, acc); sqr_sos(acc, acc);
                if (d) mul(acc, acc, tab[d]);
            } else if (d) {
                limbs_copy<L>(acc, tab[d]);
                started = true;
            }
        }
        limbs_copy<L>(r, acc);
    }
};

}  // namespace ab200
