"""GPU parity of the G2 MSM (`SWCurveConfig::msm` for bls12_381::g2::Config, curves/bls12_381/src/curves/g2.rs:54):
the same pipeline as G1 instantiated over Fq2 coordinates, compared after into_affine(), limb-exact, against the reference's own
i*G2 table (g2_uncompressed_valid_test_vectors.dat via tests/golden/), the math-level oracle (naive sum) and the identity
MSM(b_i*G2, s_i) = (sum s_i*b_i mod r)*G2.  Fq2 arithmetic itself is checked element-wise (b200_fp_op_dev field 4) against the
oracle and the reference's Fq2 KATs (curves/bls12_381/src/fields/tests.rs:1231-1391)."""
import ctypes
import json
import os
import random

import numpy as np
import pytest

import algebra_b200 as ab
from algebra_b200 import _lib
from algebra_b200 import variable_base as M
from oracle import pyoracle as O

from gpu_util import dev_empty, from_dev, stream, to_dev

pytestmark = pytest.mark.gpu
G2 = O.BLS12_381_G2
CID = 2


def fq2_encode(vals):
    fq = O.BLS12_381_FQ
    out = np.zeros((len(vals), 12), dtype=np.uint64)
    for i, (c0, c1) in enumerate(vals):
        out[i, :6] = fq.limbs(fq.to_mont(c0))
        out[i, 6:] = fq.limbs(fq.to_mont(c1))
    return out


def fq2_decode(arr):
    fq = O.BLS12_381_FQ
    return [(fq.from_mont(fq.from_limbs(r[:6])), fq.from_mont(fq.from_limbs(r[6:]))) for r in np.asarray(arr).reshape(-1, 12)]


def fp2_op(op, a, b=None):
    da, db = to_dev(a), (to_dev(b) if b is not None else None)
    out = dev_empty(da.shape)
    _lib.check(_lib.lib().b200_fp_op_dev(4, op, da.data_ptr(), db.data_ptr() if db is not None else None, out.data_ptr(), a.shape[0], 1, stream()))
    return from_dev(out)


def test_fq2_ops_vs_oracle_and_reference_kats(golden_dir):
    F, p = G2.F, O.BLS12_381_FQ.p
    rnd = random.Random(11)
    kats = json.load(open(os.path.join(golden_dir, "kats.json")))["bls12_381_fq_field"]
    A = [(rnd.randrange(p), rnd.randrange(p)) for _ in range(200)] + [(0, 0), (1, 0), (0, 1), (p - 1, p - 1), (p - 1, 0)]
    B = [(rnd.randrange(p), rnd.randrange(p)) for _ in range(200)] + [(p - 1, p - 1), (0, 0), (0, 1), (p - 1, 1), (0, p - 1)]
    ka = tuple(int(x, 16) for x in kats["fq2_mul"]["a"])
    kb = tuple(int(x, 16) for x in kats["fq2_mul"]["b"])
    A.append(ka); B.append(kb)
    sa = tuple(int(x, 16) for x in kats["fq2_square"]["a"])
    A.append(sa); B.append(sa)
    ea, eb = fq2_encode(A), fq2_encode(B)
    assert fq2_decode(fp2_op(0, ea, eb)) == [F.mul(a, b) for a, b in zip(A, B)]
    assert fq2_decode(fp2_op(1, ea, eb)) == [F.add(a, b) for a, b in zip(A, B)]
    assert fq2_decode(fp2_op(2, ea, eb)) == [F.sub(a, b) for a, b in zip(A, B)]
    assert fq2_decode(fp2_op(3, ea)) == [F.sqr(a) for a in A]
    assert fq2_decode(fp2_op(4, ea)) == [F.add(a, a) for a in A]
    assert fq2_decode(fp2_op(5, ea)) == [F.neg(a) for a in A]
    nz = [a for a in A if a != (0, 0)]
    assert fq2_decode(fp2_op(8, fq2_encode(nz))) == [F.inv(a) for a in nz]
    # the reference's own known answers
    assert fq2_decode(fp2_op(0, fq2_encode([ka]), fq2_encode([kb])))[0] == tuple(int(x, 16) for x in kats["fq2_mul"]["r"])
    assert fq2_decode(fp2_op(3, fq2_encode([sa])))[0] == tuple(int(x, 16) for x in kats["fq2_square"]["r"])
    ia = tuple(int(x, 16) for x in kats["fq2_inverse"]["a"])
    assert fq2_decode(fp2_op(8, fq2_encode([ia])))[0] == tuple(int(x, 16) for x in kats["fq2_inverse"]["r"])


def g2_ec_op(op, a, b, wo):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    da = to_dev(a)
    db = to_dev(np.ascontiguousarray(b, dtype=np.uint64)) if b is not None else None
    out = dev_empty((a.shape[0], wo * 12))
    _lib.check(_lib.lib().b200_ec_op_dev(CID, op, da.data_ptr(), db.data_ptr() if db is not None else None, out.data_ptr(), a.shape[0], stream()))
    return from_dev(out)


def test_g2_point_ops_vs_oracle():
    """ec.cuh over Fp2 on the device (b200_ec_op_dev, curve id 2): bucket +/-= affine incl. the exceptional branches, bucket += bucket,
    doubling, conversions, Jacobian add / double — compared as affine points with the oracle."""
    fq = O.BLS12_381_FQ
    rnd = random.Random(77)
    P = [G2.mul(G2.G, rnd.randrange(1, 1 << 40)) for _ in range(12)]
    Q = [G2.mul(G2.G, rnd.randrange(1, 1 << 40)) for _ in range(12)]
    Q[3] = P[3]; Q[4] = G2.neg(P[4]); Q[5] = None

    def enc2(c):
        return np.array(fq.limbs(fq.to_mont(c[0])) + fq.limbs(fq.to_mont(c[1])), dtype=np.uint64)

    def xyzz(Pt):
        if Pt is None:
            return np.concatenate([enc2((1, 0)), enc2((1, 0)), enc2((0, 0)), enc2((0, 0))])
        return np.concatenate([enc2(Pt[0]), enc2(Pt[1]), enc2((1, 0)), enc2((1, 0))])

    def aff(rows):
        return G2.decode_affine(g2_ec_op(5, g2_ec_op(4, rows, None, 3), None, 2))

    B = np.stack([xyzz(p) for p in P])
    Aq = G2.encode_affine(Q)
    assert aff(g2_ec_op(0, B, Aq, 4)) == [G2.add(p, q) for p, q in zip(P, Q)]
    assert aff(g2_ec_op(1, B, Aq, 4)) == [G2.add(p, G2.neg(q)) for p, q in zip(P, Q)]
    assert aff(g2_ec_op(0, np.stack([xyzz(None)] * len(Q)), Aq, 4)) == Q
    S = g2_ec_op(0, B, G2.encode_affine(P[1:] + P[:1]), 4)
    T = g2_ec_op(0, np.stack([xyzz(q) for q in Q]), G2.encode_affine(P[2:] + P[:2]), 4)
    sa, ta = aff(S), aff(T)
    assert aff(g2_ec_op(2, S, T, 4)) == [G2.add(x, y) for x, y in zip(sa, ta)]
    assert aff(g2_ec_op(2, S, S, 4)) == [G2.add(x, x) for x in sa]
    assert aff(g2_ec_op(3, S, None, 4)) == [G2.add(x, x) for x in sa]
    J, K = g2_ec_op(4, S, None, 3), g2_ec_op(4, T, None, 3)
    assert G2.decode_affine(g2_ec_op(5, g2_ec_op(6, J, K, 3), None, 2)) == [G2.add(x, y) for x, y in zip(sa, ta)]
    assert G2.decode_affine(g2_ec_op(5, g2_ec_op(7, J, None, 3), None, 2)) == [G2.add(x, x) for x in sa]


def test_g2_tiny_msms():
    """n = 1, 2, 3 with scalars 1, 2, r-1: isolates the MSM plumbing (digits, buckets, reduction, combine) from the volume"""
    fr = G2.fr
    P5, P9 = G2.mul(G2.G, 5), G2.mul(G2.G, 9)
    for pts, sc in (([P5], [1]), ([P5], [2]), ([P5], [fr.p - 1]), ([P5, P9], [1, 1]), ([P5, P9], [3, fr.p - 2]), ([P5, None, P9], [7, 5, 0])):
        want = G2.encode_affine([G2.naive_msm(pts, sc)])[0]
        got = gpu_affine(G2.encode_affine(pts), fr.encode(sc))
        assert (got == want).all(), (sc,)


@pytest.fixture(scope="module")
def g2_table(golden_dir):
    return np.load(os.path.join(golden_dir, "bls12_381_g2_multiples.npy"))


def table_points(tab, idx):
    pts = []
    for i in idx:
        v = [sum(int(tab[i, 6 * k + j]) << (64 * j) for j in range(6)) for k in range(4)]
        pts.append(None if not any(v) else ((v[0], v[1]), (v[2], v[3])))
    return pts


def gpu_affine(bases, scalars, device=False):
    xyz = ab.msm(CID, to_dev(bases), to_dev(scalars)) if device else ab.msm(CID, bases, scalars)
    return ab.into_affine(CID, xyz)


def test_kat_from_reference_g2_table(g2_table):
    fr = O.BLS12_381_FR
    rnd = random.Random(7)
    for trial in range(5):
        k = rnd.randrange(1, 30)
        idx = [rnd.randrange(0, 60) for _ in range(k)]          # includes 0*G2 = identity and repeats
        sc = [rnd.randrange(0, 6) for _ in range(k)]
        total = sum(i * s for i, s in zip(idx, sc))
        if total >= 1000:
            continue
        bases = G2.encode_affine(table_points(g2_table, idx))
        want = G2.encode_affine(table_points(g2_table, [total]))[0]
        for c in (0, 3, 7):
            M.set_window(c)
            assert (gpu_affine(bases, fr.encode(sc)) == want).all(), (trial, c)
        M.set_window(0)
    bases = G2.encode_affine(table_points(g2_table, [10, 30, 7, 7]))
    want = G2.encode_affine(table_points(g2_table, [50]))[0]
    assert (gpu_affine(bases, fr.encode([fr.p - 1, 2, 5, fr.p - 5])) == want).all()      # negative classes, cancellation
    two = G2.encode_affine(table_points(g2_table, [7, 7]))
    assert (gpu_affine(two, fr.encode([5, fr.p - 5])) == 0).all()                        # -> identity
    xyz = ab.msm(CID, two, fr.encode([5, fr.p - 5]))
    fq = O.BLS12_381_FQ
    assert fq.from_limbs(xyz[:6]) == fq.R and not xyz[6:12].any() and fq.from_limbs(xyz[12:18]) == fq.R and not xyz[18:].any()   # (1, 1, 0) over Fq2
    with pytest.raises(ab.LengthMismatch):
        ab.msm(CID, two, fr.encode([1]))


def synth(n, seed):
    d_bases, d_b, d_s = dev_empty((n, 24)), dev_empty((n,)), dev_empty((n, 4))
    _lib.check(_lib.lib().b200_gen_bases_dev(CID, seed, n, d_bases.data_ptr(), d_b.data_ptr(), stream()))
    _lib.check(_lib.lib().b200_gen_scalars_dev(0, seed ^ 0x55AA, n, d_s.data_ptr(), stream()))
    return d_bases, d_b, d_s


def expected_from_b(b_host, s_host):
    fr = G2.fr
    tot = sum(int(b) * s for b, s in zip(b_host, fr.decode(s_host))) % fr.p
    return G2.encode_affine([G2.mul(G2.G, tot)])[0]


def test_generator_and_random_vs_naive():
    d_bases, d_b, d_s = synth(300, 42)
    bh, bb, sh = from_dev(d_bases), from_dev(d_b), from_dev(d_s)
    pts = G2.decode_affine(bh)
    for i in list(range(8)) + [299]:
        assert pts[i] == G2.mul(G2.G, int(bb[i])), i
    assert all(G2.on_curve(P) for P in pts)
    n = 200
    want = G2.encode_affine([G2.naive_msm(pts[:n], G2.fr.decode(sh[:n]))])[0]
    assert (want == expected_from_b(bb[:n], sh[:n])).all()
    for c in (0, 5, 11):
        M.set_window(c)
        assert (gpu_affine(bh[:n], sh[:n]) == want).all(), c
        assert (gpu_affine(bh[:n], sh[:n], device=True) == want).all(), c
    M.set_window(0)
    out = np.zeros(36, dtype=np.uint64)
    _lib.check(_lib.lib().b200_msm_sw_g2(bh[:n].ctypes.data_as(ctypes.c_void_p), sh[:n].ctypes.data_as(ctypes.c_void_p), n, out.ctypes.data_as(ctypes.c_void_p)))
    assert (ab.into_affine(CID, out) == want).all()


def test_adversarial_bases_and_affine_levels():
    """identity bases, repeated base with the same scalar (doubling inside a bucket), P and -P in one bucket, heavy buckets;
    with the batched-affine levels forced on (pair-add kernels over Fq2) and off."""
    fr = G2.fr
    rnd = random.Random(3)
    n = 1 << 12
    d_bases, d_b, _ = synth(n, 77)
    bh, bb = from_dev(d_bases).copy(), from_dev(d_b).copy().astype(object)
    sc = [rnd.randrange(fr.p) for _ in range(n)]
    bh[5] = 0; bb[5] = 0
    bh[9] = bh[8]; bb[9] = bb[8]; sc[9] = sc[8]
    p8 = G2.decode_affine(bh[8:9])[0]
    bh[11] = G2.encode_affine([G2.neg(p8)])[0]; bb[11] = -int(bb[8]); sc[11] = sc[8]
    for i in range(100, 164):
        bh[i] = bh[100]; bb[i] = bb[100]; sc[i] = sc[100]
    for i in range(2000, 2600):                      # many points in very few buckets
        sc[i] = sc[2000] if i % 2 else sc[2001]
    sh = fr.encode(sc)
    tot = sum(int(b) * s for b, s in zip(bb, sc)) % fr.p
    want = G2.encode_affine([G2.mul(G2.G, tot)])[0]
    try:
        for levels, c in ((0, 0), (1, 6), (2, 7), (3, 5), (-1, 9)):
            M.set_affine_levels(levels)
            M.set_window(c)
            assert (gpu_affine(bh, sh) == want).all(), (levels, c)
            assert (gpu_affine(bh, sh, device=True) == want).all(), (levels, c)
    finally:
        M.set_affine_levels(-1)
        M.set_window(0)


@pytest.mark.parametrize("log_n", [10, 14, 18, 20])
def test_sizes_by_linear_identity(log_n):
    n = 1 << log_n
    d_bases, d_b, d_s = synth(n, 1000 + log_n)
    sh, bb = from_dev(d_s), from_dev(d_b)
    want = expected_from_b(bb, sh)
    assert (ab.into_affine(CID, ab.msm(CID, d_bases, d_s)) == want).all()
    if log_n <= 18:
        assert (ab.into_affine(CID, ab.msm(CID, from_dev(d_bases), sh)) == want).all()      # host-buffer entry
    if log_n == 14:   # streaming entry + sum of partials
        st = M.MsmStream(CID, 5000, n)
        bh = from_dev(d_bases)
        for lo in range(0, n, 5000):
            st.push(bh[lo:lo + 5000], sh[lo:lo + 5000])
        assert (ab.into_affine(CID, st.finish()) == want).all()
