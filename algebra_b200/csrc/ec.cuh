// ec.cuh — short-Weierstrass (a = 0) point arithmetic over Fp<P>: XYZZ buckets and Jacobian points.
//
// Device counterparts of ark_ec's types (ec/src/models/short_weierstrass/):
//   Bucket (XYZZ)           bucket.rs:21-30     zero = (1,1,0,0) :78-83, is_zero <=> zz==0 && zzz==0 :108-110
//   Bucket += Affine        bucket.rs:168-238   madd-2008-s incl. the exceptional cases (:175-200)
//   Bucket += &Bucket       bucket.rs:256-337   add-2008-s
//   Bucket::double_in_place bucket.rs:112-146   dbl-2008-s-1
//   Affine::double_to_bucket affine.rs:169-201  mdbl-2008-s-1
//   From<Bucket> for Projective bucket.rs:389-398
//   Projective (Jacobian) double/add  group.rs:171-221 / 450-538 ; zero = (1,1,0) :142-158
// The reference's `sum_of_products([r,-y1],[q,ppp])` is evaluated as r*q - y1*ppp (same field element).
// Affine infinity is (0,0) (ZeroFlag = (), short_weierstrass/mod.rs:224-230).
#pragma once
#include "fp.cuh"
#include "fp2.cuh"

namespace ab200 {

template <int L> struct Xyzz {
    uint32_t x[L], y[L], zz[L], zzz[L];
};
template <int L> struct Jac {
    uint32_t x[L], y[L], z[L];
};

// FT = the coordinate field's operations class: Fp<P> (G1, L = P::L words per coordinate) or Fp2<P> (G2, L = 2*P::L).
template <class FT> struct Ec {
    using F = FT;
    static constexpr int L = F::L;
    using B = Xyzz<L>;
    using J = Jac<L>;

    static AB_HD void xyzz_set_zero(B &b) {
        F::set_one(b.x);
        F::set_one(b.y);
        F::set_zero(b.zz);
        F::set_zero(b.zzz);
    }
    static AB_HD bool xyzz_is_zero(const B &b) { return limbs_is_zero<L>(b.zz) && limbs_is_zero<L>(b.zzz); }
    static AB_HD bool affine_is_zero(const uint32_t *x, const uint32_t *y) { return limbs_is_zero<L>(x) && limbs_is_zero<L>(y); }

    // affine.rs:169-201  (x,y) != infinity
    static AB_HD void mdbl(B &o, const uint32_t *px, const uint32_t *py) {
        uint32_t u[L], v[L], w[L], s[L], m[L], t[L];
        F::dbl(u, py);
        F::sqr(v, u);
        F::mul(w, u, v);
        F::mul(s, px, v);
        F::sqr(m, px);
        F::dbl(t, m);
        F::add(m, m, t);
        F::sqr(o.x, m);
        F::dbl(t, s);
        F::sub(o.x, o.x, t);
        F::sub(t, s, o.x);
        F::mul(t, m, t);
        F::mul(u, w, py);
        F::sub(o.y, t, u);
        limbs_copy<L>(o.zz, v);
        limbs_copy<L>(o.zzz, w);
    }

    // The generic (non-exceptional) madd-2008-s body: b += (x2, y2), with U2 = x2*ZZ1 and S2 = y2*ZZZ1 given.
    static AB_HD void madd_core(B &b, const uint32_t *u2, const uint32_t *s2) {
        uint32_t p[L], r[L], pp[L], ppp[L], q[L], t[L];
        F::sub(p, u2, b.x);
        F::sub(r, s2, b.y);
        F::sqr(pp, p);
        F::mul(ppp, pp, p);
        F::mul(q, b.x, pp);
        F::sqr(b.x, r);
        F::sub(b.x, b.x, ppp);
        F::dbl(t, q);
        F::sub(b.x, b.x, t);
        F::sub(q, q, b.x);
        F::mul(t, b.y, ppp);
        F::mul(q, r, q);
        F::sub(b.y, q, t);
        F::mul(b.zz, b.zz, pp);
        F::mul(b.zzz, b.zzz, ppp);
    }

    // bucket.rs:168-238 ; negate != 0 adds (x2, -y2)  (:240-244, affine.rs:303-306)
    static AB_HD void madd(B &b, const uint32_t *x2, const uint32_t *y2_in, bool negate) {
        if (affine_is_zero(x2, y2_in)) return;  // other = infinity (:175)
        uint32_t y2[L];
        F::cneg(y2, y2_in, negate);
        if (xyzz_is_zero(b)) {  // :176-182
            limbs_copy<L>(b.x, x2);
            limbs_copy<L>(b.y, y2);
            F::set_one(b.zz);
            F::set_one(b.zzz);
            return;
        }
        uint32_t u2[L], s2[L];
        F::mul(u2, x2, b.zz);
        F::mul(s2, y2, b.zzz);
        if (limbs_eq<L>(b.x, u2)) {
            if (limbs_eq<L>(b.y, s2)) mdbl(b, x2, y2);  // :193-196
            else xyzz_set_zero(b);                       // :197-200
            return;
        }
        madd_core(b, u2, s2);
    }

    // bucket.rs:112-146
#if defined(__CUDACC__) && defined(AB_EC_NOINLINE_WIDE)
    static __host__ __device__ __noinline__ void xyzz_dbl_call(B &b) { xyzz_dbl_impl(b); }
    static AB_HD void xyzz_dbl(B &b) {
        if (L > 12) xyzz_dbl_call(b);
        else xyzz_dbl_impl(b);
    }
#else
    static AB_HD void xyzz_dbl(B &b) { xyzz_dbl_impl(b); }
#endif
    static AB_HD void xyzz_dbl_impl(B &b) {
        uint32_t u[L], v[L], w[L], s[L], m[L], t[L];
        F::dbl(u, b.y);
        F::sqr(v, u);
        F::mul(w, u, v);
        F::mul(s, b.x, v);
        F::sqr(m, b.x);
        F::dbl(t, m);
        F::add(m, m, t);
        F::mul(u, w, b.y);  // W*Y1 (u is free now)
        F::sqr(b.x, m);
        F::dbl(t, s);
        F::sub(b.x, b.x, t);
        F::sub(t, s, b.x);
        F::mul(t, m, t);
        F::sub(b.y, t, u);
        F::mul(b.zz, b.zz, v);
        F::mul(b.zzz, b.zzz, w);
    }

    // bucket.rs:256-337
#if defined(__CUDACC__) && defined(AB_EC_NOINLINE_WIDE)
    static __host__ __device__ __noinline__ void xyzz_add_call(B &a, const B &o) { xyzz_add_impl(a, o); }
    static AB_HD void xyzz_add(B &a, const B &o) {
        if (L > 12) xyzz_add_call(a, o);
        else xyzz_add_impl(a, o);
    }
#else
    static AB_HD void xyzz_add(B &a, const B &o) { xyzz_add_impl(a, o); }
#endif
    static AB_HD void xyzz_add_impl(B &a, const B &o) {
        if (xyzz_is_zero(a)) { a = o; return; }
        if (xyzz_is_zero(o)) return;
        uint32_t u1[L], u2[L], s1[L], s2[L];
        F::mul(u1, a.x, o.zz);
        F::mul(u2, o.x, a.zz);
        F::mul(s1, a.y, o.zzz);
        F::mul(s2, o.y, a.zzz);
        if (limbs_eq<L>(u1, u2)) {
            if (limbs_eq<L>(s1, s2)) xyzz_dbl(a);
            else xyzz_set_zero(a);
            return;
        }
        uint32_t p[L], r[L], pp[L], ppp[L], q[L], t[L];
        F::sub(p, u2, u1);
        F::sub(r, s2, s1);
        F::sqr(pp, p);
        F::mul(ppp, pp, p);
        F::mul(q, u1, pp);
        F::sqr(a.x, r);
        F::sub(a.x, a.x, ppp);
        F::dbl(t, q);
        F::sub(a.x, a.x, t);
        F::sub(q, q, a.x);
        F::mul(t, s1, ppp);
        F::mul(q, r, q);
        F::sub(a.y, q, t);
        F::mul(a.zz, a.zz, pp);
        F::mul(a.zz, a.zz, o.zz);
        F::mul(a.zzz, a.zzz, ppp);
        F::mul(a.zzz, a.zzz, o.zzz);
    }

    static AB_HD void jac_set_zero(J &p) {
        F::set_one(p.x);
        F::set_one(p.y);
        F::set_zero(p.z);
    }
    static AB_HD bool jac_is_zero(const J &p) { return limbs_is_zero<L>(p.z); }

    // bucket.rs:389-398
    static AB_HD void xyzz_to_jac(J &o, const B &b) {
        if (xyzz_is_zero(b)) { jac_set_zero(o); return; }
        F::mul(o.x, b.x, b.zz);
        F::mul(o.y, b.y, b.zzz);
        limbs_copy<L>(o.z, b.zz);
    }

    // group.rs:171-221 (a = 0, base field of extension degree 1)
    static AB_HD void jac_dbl(J &p) {
        if (jac_is_zero(p)) return;
        uint32_t a[L], b[L], c[L], d[L], e[L], t[L];
        F::sqr(a, p.x);
        F::sqr(b, p.y);
        F::sqr(c, b);
        F::mul(d, p.x, b);
        F::dbl(d, d);
        F::dbl(d, d);
        F::dbl(t, a);
        F::add(e, a, t);
        F::mul(p.z, p.z, p.y);
        F::dbl(p.z, p.z);
        F::sqr(p.x, e);
        F::dbl(t, d);
        F::sub(p.x, p.x, t);
        F::sub(t, d, p.x);
        F::mul(p.y, t, e);
        F::dbl(c, c);
        F::dbl(c, c);
        F::dbl(c, c);
        F::sub(p.y, p.y, c);
    }

    // group.rs:450-538 (add-2007-bl)
    static AB_HD void jac_add(J &s, const J &o) {
        if (jac_is_zero(s)) { s = o; return; }
        if (jac_is_zero(o)) return;
        uint32_t z1z1[L], z2z2[L], u1[L], u2[L], s1[L], s2[L];
        F::sqr(z1z1, s.z);
        F::sqr(z2z2, o.z);
        F::mul(u1, s.x, z2z2);
        F::mul(u2, o.x, z1z1);
        F::mul(s1, s.y, o.z);
        F::mul(s1, s1, z2z2);
        F::mul(s2, o.y, s.z);
        F::mul(s2, s2, z1z1);
        if (limbs_eq<L>(u1, u2)) {
            if (limbs_eq<L>(s1, s2)) jac_dbl(s);
            else jac_set_zero(s);
            return;
        }
        uint32_t h[L], i[L], j[L], r[L], v[L], t[L];
        F::sub(h, u2, u1);
        F::dbl(i, h);
        F::sqr(i, i);
        F::neg(j, h);
        F::mul(j, j, i);
        F::sub(r, s2, s1);
        F::dbl(r, r);
        F::mul(v, u1, i);
        F::sqr(s.x, r);
        F::add(s.x, s.x, j);
        F::dbl(t, v);
        F::sub(s.x, s.x, t);
        F::sub(v, v, s.x);
        F::dbl(s1, s1);
        F::mul(s1, s1, j);
        F::mul(t, r, v);
        F::add(s.y, t, s1);
        F::mul(s.z, s.z, o.z);
        F::dbl(s.z, s.z);
        F::mul(s.z, s.z, h);
    }

    // Jacobian -> affine (affine.rs:374-396); infinity -> (0,0)
    static AB_HD void jac_to_affine(uint32_t *ax, uint32_t *ay, const J &p) {
        if (jac_is_zero(p)) { F::set_zero(ax); F::set_zero(ay); return; }
        uint32_t zi[L], zi2[L];
        F::inv(zi, p.z);
        F::sqr(zi2, zi);
        F::mul(ax, p.x, zi2);
        F::mul(ay, p.y, zi2);
        F::mul(ay, ay, zi);
    }
};

}  // namespace ab200
