// Host-side exerciser of the *device* arithmetic headers (fp.cuh / ec.cuh) through their PTX emulation
// path, so the algorithm text that runs on the GPU can be checked against the oracle on the CPU-only
// build box (tests/test_host_selftest.py).  Test tool: not part of the product library.
#include "../algebra_b200/csrc/ec.cuh"
#include <cstring>
using namespace ab200;

template <class P> static int fp_op(int op, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n) {
    constexpr int L = P::L;
    using F = Fp<P>;
    for (size_t i = 0; i < n; i++) {
        const uint32_t *x = a + i * L, *y = b + i * L;
        uint32_t *r = out + i * L;
        switch (op) {
            case 0: F::mul(r, x, y); break;
            case 1: F::add(r, x, y); break;
            case 2: F::sub(r, x, y); break;
            case 3: F::sqr(r, x); break;
            case 4: F::dbl(r, x); break;
            case 5: F::neg(r, x); break;
            case 6: F::from_mont(r, x); break;
            case 7: F::to_mont(r, x); break;
            case 8: F::inv(r, x); break;
            case 9: F::sqr_sos(r, x); break;
            case 10: F::inv_lowlat(r, x); break;
            default: return 2;
        }
    }
    return 0;
}
template <class FT> static int ec_op(int op, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n) {
    constexpr int L = FT::L;
    using E = Ec<FT>;
    for (size_t i = 0; i < n; i++) {
        typename E::B x, y; typename E::J j, k;
        switch (op) {
            case 0: case 1: memcpy(&x, a + i * 4 * L, sizeof x); E::madd(x, b + i * 2 * L, b + i * 2 * L + L, op == 1); memcpy(out + i * 4 * L, &x, sizeof x); break;
            case 2: memcpy(&x, a + i * 4 * L, sizeof x); memcpy(&y, b + i * 4 * L, sizeof y); E::xyzz_add(x, y); memcpy(out + i * 4 * L, &x, sizeof x); break;
            case 3: memcpy(&x, a + i * 4 * L, sizeof x); E::xyzz_dbl(x); memcpy(out + i * 4 * L, &x, sizeof x); break;
            case 4: memcpy(&x, a + i * 4 * L, sizeof x); E::xyzz_to_jac(j, x); memcpy(out + i * 3 * L, &j, sizeof j); break;
            case 5: memcpy(&j, a + i * 3 * L, sizeof j); E::jac_to_affine(out + i * 2 * L, out + i * 2 * L + L, j); break;
            case 6: memcpy(&j, a + i * 3 * L, sizeof j); memcpy(&k, b + i * 3 * L, sizeof k); E::jac_add(j, k); memcpy(out + i * 3 * L, &j, sizeof j); break;
            case 7: memcpy(&j, a + i * 3 * L, sizeof j); E::jac_dbl(j); memcpy(out + i * 3 * L, &j, sizeof j); break;
            default: return 2;
        }
    }
    return 0;
}
extern "C" int selftest_fp_op(int field, int op, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n) {
    switch (field) {
        case 4: return fp_op<BlsFqRolled>(op, a, b, out, n);   // rolled-row variants used by the MSM accumulation kernel
        case 5: return fp_op<BnFqRolled>(op, a, b, out, n);
        case 0: return fp_op<BlsFq>(op, a, b, out, n);
        case 1: return fp_op<BlsFr>(op, a, b, out, n);
        case 2: return fp_op<BnFq>(op, a, b, out, n);
        case 3: return fp_op<BnFr>(op, a, b, out, n);
    }
    return 1;
}
extern "C" int selftest_ec_op(int curve, int op, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n) {
    return curve == 0 ? ec_op<Fp<BlsFq>>(op, a, b, out, n) : curve == 1 ? ec_op<Fp<BnFq>>(op, a, b, out, n)
         : curve == 2 ? ec_op<Fp2<BlsFq>>(op, a, b, out, n) : 1;   // 2 = G2 of BLS12-381 (coordinates in Fq2)
}
