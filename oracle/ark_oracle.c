/*
 * ark_oracle.c — limb-level CPU restatement (O2) of the arkworks-rs/algebra v0.6.0
 * hot path: Montgomery Fp, short-Weierstrass XYZZ/Jacobian formulas, signed-digit
 * Pippenger MSM, radix-2 DIF/DIT NTT.
 *
 * TEST INFRASTRUCTURE ONLY.  The product (algebra_b200/, include/) never links,
 * loads or calls this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do, as the checker / the timed CPU stand-in.
 *
 * It is a restatement, not a copy: every routine cites the reference file:line whose
 * algorithm it follows (paths relative to /root/reference).  It is pinned against the
 * reference's golden vectors (i*G table, Montgomery constants, Fq2 KATs) and against
 * the independent Python big-int oracle by tests/test_oracle_golden.py.
 *
 * The Rust reference cannot be built in this image (no cargo/rustc, deps not vendored),
 * so this file — compiled `gcc -O3 -march=native -pthread` — is also the "port"-kind CPU
 * baseline.  Deviations from the reference's threading, both result-neutral:
 *   - rayon's (chunk x 2-thread pool) nesting is flattened into one pthread work queue over
 *     (chunk, window) tasks; chunking and the per-chunk window size c are kept
 *     (ec/src/scalar_mul/variable_base/mod.rs:445-449,521-535);
 *   - msm_signed's small-scalar size classes (:251-336) are folded into the signed-digit
 *     path (zero scalars are still filtered, :253); the sum is the same group element.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>
#include <sched.h>
#include <unistd.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;
#define MAXN 6

typedef struct {
    int N;            /* u64 limbs */
    int bits;         /* MODULUS_BIT_SIZE */
    u64 p[MAXN];      /* modulus */
    u64 inv;          /* -p^-1 mod 2^64  (montgomery_backend.rs:520-538) */
    u64 R[MAXN];      /* 2^(64N) mod p = ONE */
    u64 R2[MAXN];     /* R^2 mod p */
    int ready;
} field_t;

/* moduli: curves/bls12_381/src/fields/{fq,fr}.rs, curves/bn254/src/fields/{fq,fr}.rs */
static field_t F_BLS_FQ = {6, 381, {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                                     0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL}};
static field_t F_BLS_FR = {4, 255, {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL,
                                     0x73eda753299d7d48ULL}};
static field_t F_BN_FQ = {4, 254, {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL,
                                    0x30644e72e131a029ULL}};
static field_t F_BN_FR = {4, 254, {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL,
                                    0x30644e72e131a029ULL}};

/* ---------------- BigInt<N> primitives (ff/src/biginteger/arithmetic.rs:6-113) ---------------- */
static inline u64 adc(u64 a, u64 b, u64 *carry) { u128 t = (u128)a + b + *carry; *carry = (u64)(t >> 64); return (u64)t; }
static inline u64 sbb(u64 a, u64 b, u64 *borrow) { u128 t = (u128)a - b - *borrow; *borrow = (u64)(t >> 64) & 1; return (u64)t; }
static inline u64 mac_with_carry(u64 a, u64 b, u64 c, u64 *carry) { u128 t = (u128)b * c + a + *carry; *carry = (u64)(t >> 64); return (u64)t; }

static inline int geq(const u64 *a, const u64 *b, int N) { /* Ord: MS limb first, biginteger/mod.rs:593-616 */
    for (int i = N - 1; i >= 0; i--) { if (a[i] != b[i]) return a[i] > b[i]; }
    return 1;
}
static inline int is_zero(const u64 *a, int N) { u64 o = 0; for (int i = 0; i < N; i++) o |= a[i]; return o == 0; }
static inline int eq(const u64 *a, const u64 *b, int N) { u64 o = 0; for (int i = 0; i < N; i++) o |= a[i] ^ b[i]; return o == 0; }
static inline void add_nocarry(u64 *a, const u64 *b, int N) { u64 c = 0; for (int i = 0; i < N; i++) a[i] = adc(a[i], b[i], &c); }
static inline void sub_noborrow(u64 *a, const u64 *b, int N) { u64 c = 0; for (int i = 0; i < N; i++) a[i] = sbb(a[i], b[i], &c); }

/* ---------------- Fp ops (ff/src/fields/models/fp/montgomery_backend.rs) ---------------- */
#define FP_INLINE static inline __attribute__((always_inline))

FP_INLINE void fp_reduce(const field_t *f, u64 *a) { /* subtract_modulus, fp/mod.rs:140-155 */
    if (geq(a, f->p, f->N)) sub_noborrow(a, f->p, f->N);
}
FP_INLINE void fp_add(const field_t *f, u64 *a, const u64 *b) { add_nocarry(a, b, f->N); fp_reduce(f, a); }        /* :128-138 */
FP_INLINE void fp_sub(const field_t *f, u64 *a, const u64 *b) {                                                    /* :141-148 */
    if (!geq(a, b, f->N)) add_nocarry(a, f->p, f->N);
    sub_noborrow(a, b, f->N);
}
FP_INLINE void fp_dbl(const field_t *f, u64 *a) {                                                                  /* :151-161 */
    u64 c = 0; for (int i = 0; i < f->N; i++) { u64 t = a[i]; a[i] = (t << 1) | c; c = t >> 63; }
    fp_reduce(f, a);
}
FP_INLINE void fp_neg(const field_t *f, u64 *a) {                                                                  /* :164-171 */
    if (!is_zero(a, f->N)) { u64 t[MAXN]; memcpy(t, f->p, 8 * f->N); sub_noborrow(t, a, f->N); memcpy(a, t, 8 * f->N); }
}
/* CIOS with the no-carry optimisation, :214-233 */
#define DEFINE_MUL(NN)                                                                         \
    FP_INLINE void fp_mul_##NN(const field_t *f, u64 *a, const u64 *b) {                       \
        u64 r[NN] = {0};                                                                       \
        for (int i = 0; i < NN; i++) {                                                         \
            u64 c1 = 0, c2 = 0;                                                                \
            r[0] = mac_with_carry(r[0], a[0], b[i], &c1);                                      \
            u64 k = r[0] * f->inv;                                                             \
            (void)mac_with_carry(r[0], k, f->p[0], &c2);                                       \
            for (int j = 1; j < NN; j++) {                                                     \
                r[j] = mac_with_carry(r[j], a[j], b[i], &c1);                                  \
                r[j - 1] = mac_with_carry(r[j], k, f->p[j], &c2);                              \
            }                                                                                  \
            r[NN - 1] = c1 + c2;                                                               \
        }                                                                                      \
        memcpy(a, r, 8 * NN);                                                                  \
        fp_reduce(f, a);                                                                       \
    }
DEFINE_MUL(4)
DEFINE_MUL(6)
FP_INLINE void fp_mul(const field_t *f, u64 *a, const u64 *b) { if (f->N == 4) fp_mul_4(f, a, b); else fp_mul_6(f, a, b); }
FP_INLINE void fp_sqr(const field_t *f, u64 *a) { u64 t[MAXN]; memcpy(t, a, 8 * f->N); fp_mul(f, a, t); }  /* asm path: mul(a,a), ff-asm/src/lib.rs:132 */

static void fp_into_bigint(const field_t *f, u64 *r) { /* :392-412 */
    int N = f->N;
    for (int i = 0; i < N; i++) {
        u64 k = r[i] * f->inv, carry = 0;
        (void)mac_with_carry(r[i], k, f->p[0], &carry);
        for (int j = 1; j < N; j++) r[(j + i) % N] = mac_with_carry(r[(j + i) % N], k, f->p[j], &carry);
        r[i] = carry;
    }
}
static void fp_from_bigint(const field_t *f, u64 *r) { if (!is_zero(r, f->N)) fp_mul(f, r, f->R2); } /* :380-390 */

static void field_init(field_t *f) {
    if (f->ready) return;
    int N = f->N;
    u64 inv = 1; /* Newton: inv = p^-1 mod 2^64, then negate (:520-538 computes the same value) */
    for (int i = 0; i < 63; i++) { inv = inv * inv; inv = inv * f->p[0]; }
    f->inv = (u64)0 - inv;
    /* R = 2^(64N) mod p by 64N modular doublings of 1; R2 by 64N more doublings of R */
    u64 t[MAXN] = {1};
    for (int i = 0; i < 64 * N; i++) fp_dbl(f, t);
    memcpy(f->R, t, 8 * N);
    for (int i = 0; i < 64 * N; i++) fp_dbl(f, t);
    memcpy(f->R2, t, 8 * N);
    f->ready = 1;
}
static void fp_pow(const field_t *f, u64 *out, const u64 *base, const u64 *exp, int explimbs) {
    u64 acc[MAXN]; memcpy(acc, f->R, 8 * f->N);
    for (int i = explimbs * 64 - 1; i >= 0; i--) {
        fp_sqr(f, acc);
        if ((exp[i / 64] >> (i % 64)) & 1) fp_mul(f, acc, base);
    }
    memcpy(out, acc, 8 * f->N);
}
static void fp_inv(const field_t *f, u64 *a) { /* a^(p-2); same value as the reference's BEA inverse (:319-378) */
    u64 e[MAXN]; memcpy(e, f->p, 8 * f->N);
    u64 two[MAXN] = {2}; sub_noborrow(e, two, f->N);
    u64 b[MAXN]; memcpy(b, a, 8 * f->N);
    fp_pow(f, a, b, e, f->N);
}

static field_t *get_field(int id) {
    field_t *f = id == 0 ? &F_BLS_FQ : id == 1 ? &F_BLS_FR : id == 2 ? &F_BN_FQ : id == 3 ? &F_BN_FR : NULL;
    if (f) field_init(f);
    return f;
}

/* field ids: 0 = BLS12-381 Fq, 1 = BLS12-381 Fr, 2 = BN254 Fq, 3 = BN254 Fr.
 * op: 0 mul, 1 add, 2 sub, 3 sqr(a), 4 dbl(a), 5 neg(a), 6 into_bigint(a), 7 from_bigint(a), 8 inverse(a) */
int ark_fp_op(int field, int op, const u64 *a, const u64 *b, u64 *out, size_t n) {
    field_t *f = get_field(field);
    if (!f) return 1;
    int N = f->N;
    for (size_t i = 0; i < n; i++) {
        u64 t[MAXN]; memcpy(t, a + i * N, 8 * N);
        switch (op) {
            case 0: fp_mul(f, t, b + i * N); break;
            case 1: fp_add(f, t, b + i * N); break;
            case 2: fp_sub(f, t, b + i * N); break;
            case 3: fp_sqr(f, t); break;
            case 4: fp_dbl(f, t); break;
            case 5: fp_neg(f, t); break;
            case 6: fp_into_bigint(f, t); break;
            case 7: fp_from_bigint(f, t); break;
            case 8: fp_inv(f, t); break;
            default: return 2;
        }
        memcpy(out + i * N, t, 8 * N);
    }
    return 0;
}
int ark_field_constants(int field, u64 *p, u64 *R, u64 *R2, u64 *inv) {
    field_t *f = get_field(field);
    if (!f) return 1;
    memcpy(p, f->p, 8 * f->N); memcpy(R, f->R, 8 * f->N); memcpy(R2, f->R2, 8 * f->N); *inv = f->inv;
    return 0;
}

/* ---------------- curves ---------------- */
typedef struct { const field_t *fq; const field_t *fr; int N; } curve_t;
static int get_curve(int id, curve_t *c) {
    if (id == 0) { c->fq = get_field(0); c->fr = get_field(1); }
    else if (id == 1) { c->fq = get_field(2); c->fr = get_field(3); }
    else return 1;
    c->N = c->fq->N;
    return 0;
}

/* Bucket (XYZZ) = x,y,zz,zzz ; zero = (1,1,0,0) bucket.rs:78-83 ; is_zero :108-110 */
typedef struct { u64 x[MAXN], y[MAXN], zz[MAXN], zzz[MAXN]; } xyzz_t;
typedef struct { u64 x[MAXN], y[MAXN], z[MAXN]; } jac_t;

static void xyzz_set_zero(const field_t *f, xyzz_t *b) {
    memset(b, 0, sizeof *b); memcpy(b->x, f->R, 8 * f->N); memcpy(b->y, f->R, 8 * f->N);
}
static inline int xyzz_is_zero(const field_t *f, const xyzz_t *b) { return is_zero(b->zz, f->N) && is_zero(b->zzz, f->N); }

/* affine.rs:169-201  mdbl-2008-s-1, a=0 */
static void affine_double_to_bucket(const field_t *f, xyzz_t *o, const u64 *px, const u64 *py) {
    int N = f->N; size_t B = 8 * N;
    u64 u[MAXN], v[MAXN], w[MAXN], s[MAXN], m[MAXN], t[MAXN], x3[MAXN], y3[MAXN];
    memcpy(u, py, B); fp_dbl(f, u);
    memcpy(v, u, B); fp_sqr(f, v);
    memcpy(w, u, B); fp_mul(f, w, v);
    memcpy(s, px, B); fp_mul(f, s, v);
    memcpy(m, px, B); fp_sqr(f, m); memcpy(t, m, B); fp_dbl(f, t); fp_add(f, m, t);
    memcpy(x3, m, B); fp_sqr(f, x3); memcpy(t, s, B); fp_dbl(f, t); fp_sub(f, x3, t);
    memcpy(y3, s, B); fp_sub(f, y3, x3); fp_mul(f, y3, m);
    memcpy(t, w, B); fp_mul(f, t, py); fp_sub(f, y3, t);
    memcpy(o->x, x3, B); memcpy(o->y, y3, B); memcpy(o->zz, v, B); memcpy(o->zzz, w, B);
}

/* bucket.rs:168-238  Bucket += Affine (madd-2008-s); neg!=0 adds (x,-y) (:240-244, affine.rs:303-306) */
static void xyzz_madd(const field_t *f, xyzz_t *b, const u64 *px, const u64 *py_in, int neg) {
    int N = f->N; size_t B = 8 * N;
    if (is_zero(px, N) && is_zero(py_in, N)) return;             /* other = infinity (:175) */
    u64 py[MAXN]; memcpy(py, py_in, B); if (neg) fp_neg(f, py);
    if (xyzz_is_zero(f, b)) {                                   /* :176-182 */
        memcpy(b->x, px, B); memcpy(b->y, py, B); memcpy(b->zz, f->R, B); memcpy(b->zzz, f->R, B);
        return;
    }
    u64 u2[MAXN], s2[MAXN];
    memcpy(u2, px, B); fp_mul(f, u2, b->zz);
    memcpy(s2, py, B); fp_mul(f, s2, b->zzz);
    if (eq(b->x, u2, N)) {
        if (eq(b->y, s2, N)) affine_double_to_bucket(f, b, px, py);   /* :193-196 */
        else xyzz_set_zero(f, b);                                      /* :197-200 */
        return;
    }
    u64 p[MAXN], r[MAXN], pp[MAXN], ppp[MAXN], q[MAXN], t[MAXN], x3[MAXN], y3[MAXN];
    memcpy(p, u2, B); fp_sub(f, p, b->x);
    memcpy(r, s2, B); fp_sub(f, r, b->y);
    memcpy(pp, p, B); fp_sqr(f, pp);
    memcpy(ppp, pp, B); fp_mul(f, ppp, p);
    memcpy(q, b->x, B); fp_mul(f, q, pp);
    memcpy(x3, r, B); fp_sqr(f, x3); fp_sub(f, x3, ppp); memcpy(t, q, B); fp_dbl(f, t); fp_sub(f, x3, t);
    fp_sub(f, q, x3);
    memcpy(y3, r, B); fp_mul(f, y3, q); memcpy(t, b->y, B); fp_mul(f, t, ppp); fp_sub(f, y3, t); /* sum_of_products([r,-y1],[q,ppp]) */
    memcpy(b->x, x3, B); memcpy(b->y, y3, B);
    fp_mul(f, b->zz, pp); fp_mul(f, b->zzz, ppp);
}

/* bucket.rs:112-146  dbl-2008-s-1 (a = 0) */
static void xyzz_dbl(const field_t *f, xyzz_t *b) {
    int N = f->N; size_t B = 8 * N; (void)N;
    u64 u[MAXN], v[MAXN], w[MAXN], s[MAXN], m[MAXN], t[MAXN], x3[MAXN], y3[MAXN];
    memcpy(u, b->y, B); fp_dbl(f, u);
    memcpy(v, u, B); fp_sqr(f, v);
    memcpy(w, u, B); fp_mul(f, w, v);
    memcpy(s, b->x, B); fp_mul(f, s, v);
    memcpy(m, b->x, B); fp_sqr(f, m); memcpy(t, m, B); fp_dbl(f, t); fp_add(f, m, t);
    memcpy(x3, m, B); fp_sqr(f, x3); memcpy(t, s, B); fp_dbl(f, t); fp_sub(f, x3, t);
    memcpy(y3, s, B); fp_sub(f, y3, x3); fp_mul(f, y3, m);
    memcpy(t, w, B); fp_mul(f, t, b->y); fp_sub(f, y3, t);
    memcpy(b->x, x3, B); memcpy(b->y, y3, B); fp_mul(f, b->zz, v); fp_mul(f, b->zzz, w);
}

/* bucket.rs:256-337  Bucket += &Bucket (add-2008-s) */
static void xyzz_add(const field_t *f, xyzz_t *a, const xyzz_t *o) {
    int N = f->N; size_t B = 8 * N;
    if (xyzz_is_zero(f, a)) { *a = *o; return; }
    if (xyzz_is_zero(f, o)) return;
    u64 u1[MAXN], u2[MAXN], s1[MAXN], s2[MAXN];
    memcpy(u1, a->x, B); fp_mul(f, u1, o->zz);
    memcpy(u2, o->x, B); fp_mul(f, u2, a->zz);
    memcpy(s1, a->y, B); fp_mul(f, s1, o->zzz);
    memcpy(s2, o->y, B); fp_mul(f, s2, a->zzz);
    if (eq(u1, u2, N)) {
        if (eq(s1, s2, N)) xyzz_dbl(f, a); else xyzz_set_zero(f, a);
        return;
    }
    u64 p[MAXN], r[MAXN], pp[MAXN], ppp[MAXN], q[MAXN], t[MAXN], x3[MAXN], y3[MAXN];
    memcpy(p, u2, B); fp_sub(f, p, u1);
    memcpy(r, s2, B); fp_sub(f, r, s1);
    memcpy(pp, p, B); fp_sqr(f, pp);
    memcpy(ppp, pp, B); fp_mul(f, ppp, p);
    memcpy(q, u1, B); fp_mul(f, q, pp);
    memcpy(x3, r, B); fp_sqr(f, x3); fp_sub(f, x3, ppp); memcpy(t, q, B); fp_dbl(f, t); fp_sub(f, x3, t);
    fp_sub(f, q, x3);
    memcpy(y3, r, B); fp_mul(f, y3, q); memcpy(t, s1, B); fp_mul(f, t, ppp); fp_sub(f, y3, t);
    memcpy(a->x, x3, B); memcpy(a->y, y3, B);
    fp_mul(f, a->zz, pp); fp_mul(f, a->zz, o->zz);
    fp_mul(f, a->zzz, ppp); fp_mul(f, a->zzz, o->zzz);
}

/* Projective (Jacobian); zero = (1,1,0) group.rs:142-158 */
static void jac_set_zero(const field_t *f, jac_t *p) { memset(p, 0, sizeof *p); memcpy(p->x, f->R, 8 * f->N); memcpy(p->y, f->R, 8 * f->N); }
static inline int jac_is_zero(const field_t *f, const jac_t *p) { return is_zero(p->z, f->N); }

/* From<Bucket> for Projective, bucket.rs:389-398 */
static void xyzz_to_jac(const field_t *f, jac_t *o, const xyzz_t *b) {
    size_t B = 8 * f->N;
    if (xyzz_is_zero(f, b)) { jac_set_zero(f, o); return; }
    memcpy(o->x, b->x, B); fp_mul(f, o->x, b->zz);
    memcpy(o->y, b->y, B); fp_mul(f, o->y, b->zzz);
    memcpy(o->z, b->zz, B);
}
/* group.rs:171-221  double_in_place, a = 0, extension degree 1 branch */
static void jac_dbl(const field_t *f, jac_t *p) {
    size_t B = 8 * f->N;
    if (jac_is_zero(f, p)) return;
    u64 a[MAXN], b[MAXN], c[MAXN], d[MAXN], e[MAXN], t[MAXN];
    memcpy(a, p->x, B); fp_sqr(f, a);
    memcpy(b, p->y, B); fp_sqr(f, b);
    memcpy(c, b, B); fp_sqr(f, c);
    memcpy(d, p->x, B); fp_mul(f, d, b); fp_dbl(f, d); fp_dbl(f, d);
    memcpy(e, a, B); fp_dbl(f, a); fp_add(f, e, a);
    fp_mul(f, p->z, p->y); fp_dbl(f, p->z);
    memcpy(p->x, e, B); fp_sqr(f, p->x); memcpy(t, d, B); fp_dbl(f, t); fp_sub(f, p->x, t);
    memcpy(p->y, d, B); fp_sub(f, p->y, p->x); fp_mul(f, p->y, e);
    fp_dbl(f, c); fp_dbl(f, c); fp_dbl(f, c); fp_sub(f, p->y, c);
}
/* group.rs:450-538  add-2007-bl */
static void jac_add(const field_t *f, jac_t *s, const jac_t *o) {
    int N = f->N; size_t B = 8 * N;
    if (jac_is_zero(f, s)) { *s = *o; return; }
    if (jac_is_zero(f, o)) return;
    u64 z1z1[MAXN], z2z2[MAXN], u1[MAXN], u2[MAXN], s1[MAXN], s2[MAXN];
    memcpy(z1z1, s->z, B); fp_sqr(f, z1z1);
    memcpy(z2z2, o->z, B); fp_sqr(f, z2z2);
    memcpy(u1, s->x, B); fp_mul(f, u1, z2z2);
    memcpy(u2, o->x, B); fp_mul(f, u2, z1z1);
    memcpy(s1, s->y, B); fp_mul(f, s1, o->z); fp_mul(f, s1, z2z2);
    memcpy(s2, o->y, B); fp_mul(f, s2, s->z); fp_mul(f, s2, z1z1);
    if (eq(u1, u2, N)) {
        if (eq(s1, s2, N)) jac_dbl(f, s); else jac_set_zero(f, s);
        return;
    }
    u64 h[MAXN], i[MAXN], j[MAXN], r[MAXN], v[MAXN], t[MAXN];
    memcpy(h, u2, B); fp_sub(f, h, u1);
    memcpy(i, h, B); fp_dbl(f, i); fp_sqr(f, i);
    memcpy(j, h, B); fp_neg(f, j); fp_mul(f, j, i);
    memcpy(r, s2, B); fp_sub(f, r, s1); fp_dbl(f, r);
    memcpy(v, u1, B); fp_mul(f, v, i);
    memcpy(s->x, r, B); fp_sqr(f, s->x); fp_add(f, s->x, j); memcpy(t, v, B); fp_dbl(f, t); fp_sub(f, s->x, t);
    fp_sub(f, v, s->x);
    memcpy(s->y, s1, B); fp_dbl(f, s->y);
    /* sum_of_products([r, 2*s1],[v, j]) */
    fp_mul(f, s->y, j); memcpy(t, r, B); fp_mul(f, t, v); fp_add(f, s->y, t);
    fp_mul(f, s->z, o->z); fp_dbl(f, s->z); fp_mul(f, s->z, h);
}
/* affine.rs:374-396 */
static void jac_to_affine(const field_t *f, u64 *ax, u64 *ay, const jac_t *p) {
    int N = f->N; size_t B = 8 * N;
    if (jac_is_zero(f, p)) { memset(ax, 0, B); memset(ay, 0, B); return; }
    u64 zi[MAXN], zi2[MAXN];
    memcpy(zi, p->z, B); fp_inv(f, zi);
    memcpy(zi2, zi, B); fp_sqr(f, zi2);
    memcpy(ax, p->x, B); fp_mul(f, ax, zi2);
    memcpy(ay, p->y, B); fp_mul(f, ay, zi2); fp_mul(f, ay, zi);
}

/* exported EC ops for kernel-level parity tests.
 * op 0: out_xyzz = bucket + affine ; 1: bucket - affine ; 2: bucket + bucket2 ; 3: dbl(bucket)
 * 4: xyzz -> jacobian (out 3N) ; 5: jacobian(a) -> affine (out 2N) ; 6: jacobian a + jacobian b ; 7: jacobian dbl */
int ark_ec_op(int curve, int op, const u64 *a, const u64 *b, u64 *out, size_t n) {
    curve_t c; if (get_curve(curve, &c)) return 1;
    const field_t *f = c.fq; int N = c.N; size_t B = 8 * N;
    for (size_t i = 0; i < n; i++) {
        xyzz_t x, y; jac_t j, k;
        switch (op) {
            case 0: case 1:
                memcpy(x.x, a + i * 4 * N, B); memcpy(x.y, a + i * 4 * N + N, B); memcpy(x.zz, a + i * 4 * N + 2 * N, B); memcpy(x.zzz, a + i * 4 * N + 3 * N, B);
                xyzz_madd(f, &x, b + i * 2 * N, b + i * 2 * N + N, op == 1);
                memcpy(out + i * 4 * N, x.x, B); memcpy(out + i * 4 * N + N, x.y, B); memcpy(out + i * 4 * N + 2 * N, x.zz, B); memcpy(out + i * 4 * N + 3 * N, x.zzz, B);
                break;
            case 2: case 3:
                memcpy(x.x, a + i * 4 * N, B); memcpy(x.y, a + i * 4 * N + N, B); memcpy(x.zz, a + i * 4 * N + 2 * N, B); memcpy(x.zzz, a + i * 4 * N + 3 * N, B);
                if (op == 2) {
                    memcpy(y.x, b + i * 4 * N, B); memcpy(y.y, b + i * 4 * N + N, B); memcpy(y.zz, b + i * 4 * N + 2 * N, B); memcpy(y.zzz, b + i * 4 * N + 3 * N, B);
                    xyzz_add(f, &x, &y);
                } else xyzz_dbl(f, &x);
                memcpy(out + i * 4 * N, x.x, B); memcpy(out + i * 4 * N + N, x.y, B); memcpy(out + i * 4 * N + 2 * N, x.zz, B); memcpy(out + i * 4 * N + 3 * N, x.zzz, B);
                break;
            case 4:
                memcpy(x.x, a + i * 4 * N, B); memcpy(x.y, a + i * 4 * N + N, B); memcpy(x.zz, a + i * 4 * N + 2 * N, B); memcpy(x.zzz, a + i * 4 * N + 3 * N, B);
                xyzz_to_jac(f, &j, &x);
                memcpy(out + i * 3 * N, j.x, B); memcpy(out + i * 3 * N + N, j.y, B); memcpy(out + i * 3 * N + 2 * N, j.z, B);
                break;
            case 5:
                memcpy(j.x, a + i * 3 * N, B); memcpy(j.y, a + i * 3 * N + N, B); memcpy(j.z, a + i * 3 * N + 2 * N, B);
                jac_to_affine(f, out + i * 2 * N, out + i * 2 * N + N, &j);
                break;
            case 6: case 7:
                memcpy(j.x, a + i * 3 * N, B); memcpy(j.y, a + i * 3 * N + N, B); memcpy(j.z, a + i * 3 * N + 2 * N, B);
                if (op == 6) {
                    memcpy(k.x, b + i * 3 * N, B); memcpy(k.y, b + i * 3 * N + N, B); memcpy(k.z, b + i * 3 * N + 2 * N, B);
                    jac_add(f, &j, &k);
                } else jac_dbl(f, &j);
                memcpy(out + i * 3 * N, j.x, B); memcpy(out + i * 3 * N + N, j.y, B); memcpy(out + i * 3 * N + 2 * N, j.z, B);
                break;
            default: return 2;
        }
    }
    return 0;
}

/* ---------------- tiny pthread parallel-for (stand-in for rayon; libgomp is absent from this image) ---------------- */
typedef void (*range_fn)(size_t lo, size_t hi, void *ctx);
typedef struct { size_t n, grain; size_t next; range_fn fn; void *ctx; } par_job;
static void *par_worker(void *arg) {
    par_job *j = arg;
    for (;;) {
        size_t lo = __atomic_fetch_add(&j->next, j->grain, __ATOMIC_RELAXED);
        if (lo >= j->n) break;
        size_t hi = lo + j->grain > j->n ? j->n : lo + j->grain;
        j->fn(lo, hi, j->ctx);
    }
    return NULL;
}
static void par_for(size_t n, int threads, size_t grain, range_fn fn, void *ctx) {
    if (n == 0) return;
    if (grain == 0) grain = 1;
    if (threads <= 1 || n <= grain) { fn(0, n, ctx); return; }
    par_job j = {n, grain, 0, fn, ctx};
    pthread_t th[256]; if (threads > 256) threads = 256;
    for (int t = 1; t < threads; t++) pthread_create(&th[t], NULL, par_worker, &j);
    par_worker(&j);
    for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
}

/* ---------------- MSM ---------------- */
static int log2_ceil(size_t n) { int l = 0; if (n <= 1) return 0; n--; while (n) { l++; n >>= 1; } return l; } /* ark_std::log2 */
int ark_window_size(size_t n) { return n < 32 ? 3 : log2_ceil(n) * 69 / 100 + 2; } /* variable_base/mod.rs:445-449, scalar_mul/mod.rs:22-25 */

/* make_digits, variable_base/mod.rs:754-794 */
static void make_digits(const u64 *scalar, int w, int num_bits, int64_t *out) {
    const int L = 4;
    u64 radix = 1ULL << w, window_mask = radix - 1, carry = 0;
    int digits_count = (num_bits + w - 1) / w;
    for (int i = 0; i < digits_count; i++) {
        int bit_offset = i * w, u64_idx = bit_offset / 64, bit_idx = bit_offset % 64;
        u64 bit_buf;
        if (bit_idx < 64 - w || u64_idx == L - 1) bit_buf = scalar[u64_idx] >> bit_idx;
        else bit_buf = (scalar[u64_idx] >> bit_idx) | (scalar[1 + u64_idx] << (64 - bit_idx));
        u64 coef = carry + (bit_buf & window_mask);
        carry = (coef + radix / 2) >> w;
        int64_t digit = (int64_t)coef - (int64_t)(carry << w);
        if (i == digits_count - 1) digit += (int64_t)(carry << w);
        out[i] = digit;
    }
}
int ark_make_digits(const u64 *scalar_canonical, int w, int num_bits, int64_t *out) { make_digits(scalar_canonical, w, num_bits, out); return (num_bits + w - 1) / w; }

/* one window of msm_bigint_wnaf_parallel (:464-486) over a chunk */
static void window_sum(const curve_t *cv, const u64 *bases, const int64_t *digits, size_t n, int W, int c, int win,
                       xyzz_t *buckets, xyzz_t *res) {
    const field_t *f = cv->fq; int N = cv->N;
    size_t nb = (size_t)1 << c;
    for (size_t j = 0; j < nb; j++) xyzz_set_zero(f, &buckets[j]);
    for (size_t i = 0; i < n; i++) {
        int64_t d = digits[i * W + win];
        const u64 *px = bases + i * 2 * N, *py = px + N;
        if (d > 0) xyzz_madd(f, &buckets[d - 1], px, py, 0);
        else if (d < 0) xyzz_madd(f, &buckets[-d - 1], px, py, 1);
    }
    xyzz_t running; xyzz_set_zero(f, &running); xyzz_set_zero(f, res);
    for (size_t j = nb; j-- > 0;) { xyzz_add(f, &running, &buckets[j]); xyzz_add(f, res, &running); }
}

typedef struct {
    const curve_t *cv; const u64 *fb; const u64 *big; size_t m, chunk_size, nchunks; int num_bits;
    int *cs, *Ws; size_t *task_off; int64_t **digs; xyzz_t *wsum;
} msm_ctx;
static void msm_digits_range(size_t klo, size_t khi, void *vp) {
    msm_ctx *c = vp;
    for (size_t k = klo; k < khi; k++) {
        size_t lo = k * c->chunk_size, len = (lo + c->chunk_size <= c->m) ? c->chunk_size : c->m - lo;
        for (size_t i = 0; i < len; i++) make_digits(c->big + 4 * (lo + i), c->cs[k], c->num_bits, c->digs[k] + i * c->Ws[k]);
    }
}
static void msm_window_range(size_t tlo, size_t thi, void *vp) {
    msm_ctx *c = vp; int N = c->cv->N;
    for (size_t t = tlo; t < thi; t++) {
        size_t k = 0; while (c->task_off[k + 1] <= t) k++;
        int win = (int)(t - c->task_off[k]);
        size_t lo = k * c->chunk_size, len = (lo + c->chunk_size <= c->m) ? c->chunk_size : c->m - lo;
        xyzz_t *buckets = malloc(((size_t)1 << c->cs[k]) * sizeof(xyzz_t));
        window_sum(c->cv, c->fb + lo * 2 * N, c->digs[k], len, c->Ws[k], c->cs[k], win, buckets, &c->wsum[t]);
        free(buckets);
    }
}

typedef struct { const field_t *fr; const u64 *scalars; u64 *big; int zeros; } prelude_ctx;
static void prelude_range(size_t lo, size_t hi, void *vp) {
    prelude_ctx *c = vp; int z = 0;
    for (size_t i = lo; i < hi; i++) {
        u64 t[MAXN]; memcpy(t, c->scalars + 4 * i, 32); fp_into_bigint(c->fr, t);
        memcpy(c->big + 4 * i, t, 32); z |= is_zero(t, 4);
    }
    if (z) __atomic_store_n(&c->zeros, 1, __ATOMIC_RELAXED);
}

/* VariableBaseMSM::msm_unchecked (:59-64) -> msm_bigint_wnaf (:512-558) with `threads` rayon threads.
 * bases: n x 2N Montgomery limbs, scalars: n x 4 Montgomery Fr limbs; out: 3N Jacobian Montgomery limbs.
 * c_override > 0 forces the window size (for window sweeps); 0 = the reference's rule per chunk. */
int ark_msm(int curve, const u64 *bases, const u64 *scalars, size_t n, u64 *out, int threads, int c_override) {
    curve_t cv; if (get_curve(curve, &cv)) return 1;
    const field_t *f = cv.fq; int N = cv.N; size_t B = 8 * N;
    jac_t total; jac_set_zero(f, &total);
    if (threads < 1) threads = 1;
    /* into_bigint in parallel (cfg_iter, :60-62) + zero filter (msm_signed :253; compaction only if a zero exists) */
    u64 *big = malloc(n * 32 + 32); u64 *fb = NULL;
    prelude_ctx pc = {cv.fr, scalars, big, 0};
    par_for(n, threads, 1 << 14, prelude_range, &pc);
    size_t m = n;
    const u64 *use_bases = bases;
    if (pc.zeros) {
        fb = malloc(n * 2 * B + 16); m = 0;
        for (size_t i = 0; i < n; i++) {
            if (is_zero(big + 4 * i, 4)) continue;
            memmove(big + 4 * m, big + 4 * i, 32); memcpy(fb + m * 2 * N, bases + i * 2 * N, 2 * B); m++;
        }
        use_bases = fb;
    }
    if (m > 0) {
        size_t num_chunks = threads < 2 ? 1 : threads / 2;              /* :521-535 */
        size_t chunk_size = m / num_chunks; if (chunk_size == 0) chunk_size = m;
        size_t nchunks = (m + chunk_size - 1) / chunk_size;
        int num_bits = cv.fr->bits;
        /* per-chunk window and digit tables */
        int *cs = malloc(nchunks * sizeof(int)); int *Ws = malloc(nchunks * sizeof(int));
        size_t *task_off = malloc((nchunks + 1) * sizeof(size_t));
        int64_t **digs = malloc(nchunks * sizeof(int64_t *));
        task_off[0] = 0;
        for (size_t k = 0; k < nchunks; k++) {
            size_t lo = k * chunk_size, len = (lo + chunk_size <= m) ? chunk_size : m - lo;
            cs[k] = c_override > 0 ? c_override : ark_window_size(len);
            Ws[k] = (num_bits + cs[k] - 1) / cs[k];
            task_off[k + 1] = task_off[k] + Ws[k];
            digs[k] = malloc(len * Ws[k] * sizeof(int64_t));
        }
        size_t ntasks = task_off[nchunks];
        xyzz_t *wsum = malloc(ntasks * sizeof(xyzz_t));
        msm_ctx mc = {&cv, use_bases, big, m, chunk_size, nchunks, num_bits, cs, Ws, task_off, digs, wsum};
        par_for(nchunks, threads, 1, msm_digits_range, &mc);
        par_for(ntasks, threads, 1, msm_window_range, &mc);
        /* window combine per chunk (:489-502), then sum over chunks (:557) */
        for (size_t k = 0; k < nchunks; k++) {
            jac_t lowest, tot, s; xyzz_to_jac(f, &lowest, &wsum[task_off[k]]);
            jac_set_zero(f, &tot);
            for (int w = Ws[k] - 1; w >= 1; w--) {
                xyzz_to_jac(f, &s, &wsum[task_off[k] + w]);
                jac_add(f, &tot, &s);                                   /* Projective += &Bucket, bucket.rs:345-359 */
                for (int d = 0; d < cs[k]; d++) jac_dbl(f, &tot);
            }
            jac_add(f, &lowest, &tot);
            jac_add(f, &total, &lowest);
            free(digs[k]);
        }
        free(wsum); free(digs); free(task_off); free(cs); free(Ws);
    }
    free(big); free(fb);
    memcpy(out, total.x, B); memcpy(out + N, total.y, B); memcpy(out + 2 * N, total.z, B);
    return 0;
}

/* naive Σ s_i·P_i by double-and-add over canonical scalar bits (test-templates/src/msm.rs:8-15). out: affine 2N */
int ark_msm_naive(int curve, const u64 *bases, const u64 *scalars, size_t n, u64 *out_affine) {
    curve_t cv; if (get_curve(curve, &cv)) return 1;
    const field_t *f = cv.fq; int N = cv.N; size_t B = 8 * N;
    jac_t acc; jac_set_zero(f, &acc);
    for (size_t i = 0; i < n; i++) {
        u64 s[4]; memcpy(s, scalars + 4 * i, 32); fp_into_bigint(cv.fr, s);
        const u64 *px = bases + i * 2 * N, *py = px + N;
        if (is_zero(px, N) && is_zero(py, N)) continue;
        jac_t base, r; memcpy(base.x, px, B); memcpy(base.y, py, B); memcpy(base.z, f->R, B);
        jac_set_zero(f, &r);
        for (int b = 255; b >= 0; b--) { jac_dbl(f, &r); if ((s[b / 64] >> (b % 64)) & 1) jac_add(f, &r, &base); }
        jac_add(f, &acc, &r);
    }
    jac_to_affine(f, out_affine, out_affine + N, &acc);
    return 0;
}

/* ---------------- NTT (poly/src/domain/radix2/fft.rs) ---------------- */
static u64 bitrev64(u64 a, int log_len) { /* fft.rs:369-371 */
    u64 r = 0; for (int i = 0; i < 64; i++) { r = (r << 1) | (a & 1); a >>= 1; }
    return log_len == 0 ? 0 : r >> (64 - log_len);
}
static void derange(u64 *x, int log_len) { /* fft.rs:373-380 */
    size_t n = (size_t)1 << log_len;
    for (u64 idx = 1; idx + 1 < n; idx++) {
        u64 r = bitrev64(idx, log_len);
        if (idx < r) { u64 t[4]; memcpy(t, x + 4 * idx, 32); memcpy(x + 4 * idx, x + 4 * r, 32); memcpy(x + 4 * r, t, 32); }
    }
}
static void get_root_of_unity(const field_t *f, int log_n, int two_adicity, const u64 *two_adic_root, u64 *g) {
    memcpy(g, two_adic_root, 32);                       /* ff/src/fields/fft_friendly.rs:66-82 */
    for (int i = log_n; i < two_adicity; i++) fp_sqr(f, g);
}
static void fr_params(int field, int *two_adicity, u64 *gen_mont) {
    const field_t *f = get_field(field);
    u64 g[4] = {field == 1 ? 7u : 5u, 0, 0, 0};        /* GENERATOR: bls12_381 fr.rs:5, bn254 fr.rs:5 */
    fp_from_bigint(f, g);
    *two_adicity = field == 1 ? 32 : 28;
    /* TWO_ADIC_ROOT_OF_UNITY = GENERATOR^t, t = (p-1) >> s   (ff-macros/src/montgomery/mod.rs:44-55) */
    u64 t[4]; memcpy(t, f->p, 32); t[0] -= 1;
    int s = *two_adicity;
    for (int i = 0; i < 4; i++) t[i] = (t[i] >> s) | (i < 3 ? t[i + 1] << (64 - s) : 0);
    fp_pow(f, gen_mont, g, t, 4);
}

/* roots_of_unity: [1, g, ..., g^(n/2-1)]  (fft.rs:125-153) */
typedef struct { const field_t *f; const u64 *g; const u64 *c; u64 *x; } pow_ctx;
static void roots_range(size_t lo, size_t hi, void *vp) {
    pow_ctx *c = vp; u64 e[1] = {lo}, cur[MAXN];
    fp_pow(c->f, cur, c->g, e, 1);
    for (size_t i = lo; i < hi; i++) { memcpy(c->x + 4 * i, cur, 32); fp_mul(c->f, cur, c->g); }
}
static u64 *roots_of_unity(const field_t *f, const u64 *root, size_t half, int threads) {
    u64 *r = malloc((half ? half : 1) * 32);
    if (half == 0) return r;
    size_t blk = (half + threads - 1) / threads; if (blk < 1024) blk = 1024;
    pow_ctx c = {f, root, NULL, r};
    par_for(half, threads, blk, roots_range, &c);
    return r;
}

typedef struct { const field_t *f; u64 *x; const u64 *roots; size_t gap, step; } bf_ctx;
static void bf_io_range(size_t tlo, size_t thi, void *vp) { /* butterfly_fn_io, fft.rs:190-198 */
    bf_ctx *c = vp; const field_t *f = c->f; size_t gap = c->gap;
    for (size_t t = tlo; t < thi; t++) {
        size_t c0 = (t / gap) * 2 * gap, j = t % gap;
        u64 *lo = c->x + 4 * (c0 + j), *hi = lo + 4 * gap;
        u64 neg[MAXN]; memcpy(neg, lo, 32); fp_sub(f, neg, hi);
        fp_add(f, lo, hi);
        fp_mul_4(f, neg, c->roots + 4 * (j * c->step));
        memcpy(hi, neg, 32);
    }
}
static void bf_oi_range(size_t tlo, size_t thi, void *vp) { /* butterfly_fn_oi, fft.rs:201-210 */
    bf_ctx *c = vp; const field_t *f = c->f; size_t gap = c->gap;
    for (size_t t = tlo; t < thi; t++) {
        size_t c0 = (t / gap) * 2 * gap, j = t % gap;
        u64 *lo = c->x + 4 * (c0 + j), *hi = lo + 4 * gap;
        u64 h[MAXN]; memcpy(h, hi, 32); fp_mul_4(f, h, c->roots + 4 * (j * c->step));
        u64 neg[MAXN]; memcpy(neg, lo, 32); fp_sub(f, neg, h);
        fp_add(f, lo, h);
        memcpy(hi, neg, 32);
    }
}
/* io_helper (DIF, fft.rs:252-295, butterfly :190-198) — same data flow; the root-compaction trick (a cache
 * optimisation, :271-279) is kept as "stride access when few chunks, compacted copy when many". */
static void io_helper(const field_t *f, u64 *x, size_t n, const u64 *root, int threads) {
    u64 *roots = roots_of_unity(f, root, n / 2, threads);
    size_t nroots = n / 2; size_t step = 1; int first = 1;
    for (size_t gap = n / 2; gap > 0; gap /= 2) {
        size_t chunk = 2 * gap, num_chunks = n / chunk;
        if (num_chunks >= 128) {
            if (!first) { size_t s2 = step * 2, cnt = (nroots + s2 - 1) / s2; for (size_t i = 0; i < cnt; i++) memmove(roots + 4 * i, roots + 4 * i * s2, 32); nroots = cnt; }
            step = 1;
        } else step = num_chunks;
        first = 0;
        bf_ctx bc = {f, x, roots, gap, step};
        par_for(n / 2, threads, 4096, bf_io_range, &bc);
    }
    free(roots);
}
/* oi_helper (DIT, fft.rs:297-349, butterfly :201-210) */
static void oi_helper(const field_t *f, u64 *x, size_t n, const u64 *root, int threads) {
    u64 *roots = roots_of_unity(f, root, n / 2, threads);
    u64 *compact = malloc((n / 4 + 1) * 32);
    for (size_t gap = 1; gap < n; gap *= 2) {
        size_t chunk = 2 * gap, num_chunks = n / chunk, step; const u64 *rt;
        if (num_chunks >= 128 && gap < n / 2) {
            for (size_t i = 0; i < gap; i++) memcpy(compact + 4 * i, roots + 4 * i * num_chunks, 32);
            rt = compact; step = 1;
        } else { rt = roots; step = num_chunks; }
        bf_ctx bc = {f, x, rt, gap, step};
        par_for(n / 2, threads, 4096, bf_oi_range, &bc);
    }
    free(compact); free(roots);
}
/* distribute_powers_and_mul_by_const (poly/src/domain/mod.rs:119-148) */
static void dist_range(size_t lo, size_t hi, void *vp) {
    pow_ctx *c = vp; u64 e[1] = {lo}, pw[MAXN], v[MAXN];
    fp_pow(c->f, pw, c->g, e, 1); fp_mul(c->f, pw, c->c);
    for (size_t i = lo; i < hi; i++) { memcpy(v, c->x + 4 * i, 32); fp_mul(c->f, v, pw); memcpy(c->x + 4 * i, v, 32); fp_mul(c->f, pw, c->g); }
}
static void distribute_powers(const field_t *f, u64 *x, size_t n, const u64 *g, const u64 *c, int threads) {
    size_t per = n / threads; if (per < 1024) per = 1024;
    pow_ctx pc = {f, g, c, x};
    par_for(n, threads, per, dist_range, &pc);
}
static void scale_range(size_t lo, size_t hi, void *vp) {
    pow_ctx *c = vp; u64 v[MAXN];
    for (size_t i = lo; i < hi; i++) { memcpy(v, c->x + 4 * i, 32); fp_mul(c->f, v, c->c); memcpy(c->x + 4 * i, v, 32); }
}

/* Radix2EvaluationDomain::{fft_in_place, ifft_in_place} on exactly n = 2^log_n elements (radix2/mod.rs:140-153 after
 * the resize).  field: 1 = BLS12-381 Fr, 3 = BN254 Fr.  offset: NULL or 4 Montgomery limbs of the coset offset. */
int ark_fft(int field, u64 *data, unsigned log_n, int inverse, const u64 *offset, int threads) {
    if (field != 1 && field != 3) return 1;
    const field_t *f = get_field(field);
    int s; u64 w[4]; fr_params(field, &s, w);
    if ((int)log_n > s) return 2;                         /* radix2/mod.rs:61-63 */
    if (threads < 1) threads = 1;
    size_t n = (size_t)1 << log_n;
    u64 g[4]; get_root_of_unity(f, log_n, s, w, g);
    int coset = offset && !eq(offset, f->R, 4);
    if (!inverse) {                                         /* in_order_fft_in_place, fft.rs:74-79 */
        if (coset) distribute_powers(f, data, n, offset, f->R, threads);
        io_helper(f, data, n, g, threads);
        derange(data, log_n);
    } else {                                                /* in_order_ifft_in_place, fft.rs:81-88 */
        u64 ginv[4]; memcpy(ginv, g, 32); fp_inv(f, ginv);
        u64 ninv[4] = {n, 0, 0, 0}; fp_from_bigint(f, ninv); fp_inv(f, ninv);
        derange(data, log_n);
        oi_helper(f, data, n, ginv, threads);
        if (!coset) {
            pow_ctx pc = {f, NULL, ninv, data};
            par_for(n, threads, 4096, scale_range, &pc);
        } else {
            u64 oinv[4]; memcpy(oinv, offset, 32); fp_inv(f, oinv);
            distribute_powers(f, data, n, oinv, ninv, threads);
        }
    }
    return 0;
}

/* domain parameters for the host-side mirror's tests: group_gen, group_gen_inv, size_inv (Montgomery limbs) */
int ark_domain_params(int field, unsigned log_n, u64 *group_gen, u64 *group_gen_inv, u64 *size_inv) {
    if (field != 1 && field != 3) return 1;
    const field_t *f = get_field(field);
    int s; u64 w[4]; fr_params(field, &s, w);
    if ((int)log_n > s) return 2;
    get_root_of_unity(f, log_n, s, w, group_gen);
    memcpy(group_gen_inv, group_gen, 32); fp_inv(f, group_gen_inv);
    u64 ninv[4] = {(u64)1 << log_n, 0, 0, 0}; fp_from_bigint(f, ninv); fp_inv(f, ninv);
    memcpy(size_inv, ninv, 32);
    return 0;
}

/* usable hardware threads: affinity mask, capped by a cgroup-v2 CPU quota if one is set */
int ark_num_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) { int c = CPU_COUNT(&set); if (c > 0 && c < n) n = c; }
    FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r");
    if (f) {
        char q[64]; long period = 0;
        if (fscanf(f, "%63s %ld", q, &period) == 2 && strcmp(q, "max") != 0 && period > 0) {
            long quota = atol(q), c = (quota + period - 1) / period;
            if (c > 0 && c < n) n = c;
        }
        fclose(f);
    }
    return n < 1 ? 1 : (int)n;
}
