"""Pins both oracles (oracle/pyoracle.py = O1 math level, oracle/ark_oracle.c = O2 limb level)
against the reference's own golden vectors (tests/golden/, extracted by make_golden.py) and
against each other.  Mirrors: test-templates/src/fields.rs:509-556 (Montgomery config),
curves/bls12_381/src/curves/tests/mod.rs:70-111 (i*G table), curves/bls12_381/src/fields/tests.rs
(Fq2 KATs), test-templates/src/msm.rs:17-32 (MSM vs naive), poly/src/domain/radix2/mod.rs:351-391,
438-535 (FFT vs Horner / textbook)."""
import json
import os
import random

import numpy as np
import pytest

from oracle import coracle as C
from oracle import pyoracle as O


@pytest.fixture(scope="module")
def kats(golden_dir):
    return json.load(open(os.path.join(golden_dir, "kats.json")))


@pytest.fixture(scope="module")
def g1_table(golden_dir):
    tab = np.load(os.path.join(golden_dir, "bls12_381_g1_multiples.npy"))
    pts = []
    for row in tab:
        x = O.BLS12_381_FQ.from_limbs(row[:6])
        y = O.BLS12_381_FQ.from_limbs(row[6:])
        pts.append(None if x == 0 and y == 0 else (x, y))
    return pts


def test_montgomery_constants(kats):
    fq, fr = O.BLS12_381_FQ, O.BLS12_381_FR
    k = kats["bls12_381_fq"]
    assert hex(fq.INV) == k["INV"] and hex(fq.R) == k["R"] and hex(fq.R2) == k["R2"]
    assert hex(fq.to_mont(2)) == k["GENERATOR_MONT"]
    assert hex(fq.trace) == k["TRACE"]
    assert hex(fr.INV) == kats["bls12_381_fr"]["INV"] and hex(fr.p) == kats["bls12_381_fr"]["MODULUS"]
    assert hex(fq.to_mont(fq.p - 1)) == kats["bls12_381_fq_field"]["neg_one_mont"]
    # C oracle derives the same constants
    for fid, f in ((0, fq), (1, fr), (2, O.BN254_FQ), (3, O.BN254_FR)):
        p, R, R2, inv = C.field_constants(fid)
        assert f.from_limbs(p) == f.p and f.from_limbs(R) == f.R and f.from_limbs(R2) == f.R2 and inv == f.INV


def test_fq2_kats_pin_fq_mul(kats):
    """(a0+a1 u)(b0+b1 u) = (a0b0 - a1b1) + (a0b1 + a1b0) u, u^2 = -1: each KAT pins 4 Fq Montgomery products."""
    fq = O.BLS12_381_FQ
    f = kats["bls12_381_fq_field"]

    def c_mul(a, b):
        out = C.fp_op(0, "mul", fq.encode([a]), fq.encode([b]))
        return fq.decode(out)[0]

    def py_mul(a, b):
        return fq.from_mont(fq.mont_mul_cios(fq.to_mont(a), fq.to_mont(b)))

    for mul in (c_mul, py_mul):
        a0, a1 = (int(x, 16) for x in f["fq2_mul"]["a"])
        b0, b1 = (int(x, 16) for x in f["fq2_mul"]["b"])
        r0, r1 = (int(x, 16) for x in f["fq2_mul"]["r"])
        assert (mul(a0, b0) - mul(a1, b1)) % fq.p == r0
        assert (mul(a0, b1) + mul(a1, b0)) % fq.p == r1
        a0, a1 = (int(x, 16) for x in f["fq2_square"]["a"])
        r0, r1 = (int(x, 16) for x in f["fq2_square"]["r"])
        assert (mul(a0, a0) - mul(a1, a1)) % fq.p == r0
        assert (2 * mul(a0, a1)) % fq.p == r1
    # inverse KAT: (a0 + a1 u)^-1 = (a0 - a1 u) / (a0^2 + a1^2)
    a0, a1 = (int(x, 16) for x in f["fq2_inverse"]["a"])
    r0, r1 = (int(x, 16) for x in f["fq2_inverse"]["r"])
    norm = (a0 * a0 + a1 * a1) % fq.p
    ninv = fq.decode(C.fp_op(0, "inv", fq.encode([norm])))[0]
    assert a0 * ninv % fq.p == r0 and (-a1 * ninv) % fq.p == r1


@pytest.mark.parametrize("fid,f", [(0, O.BLS12_381_FQ), (1, O.BLS12_381_FR), (2, O.BN254_FQ), (3, O.BN254_FR)])
def test_field_ops_c_vs_python(fid, f):
    rnd = random.Random(1234 + fid)
    edge = [0, 1, 2, f.p - 1, f.p - 2, f.R, f.p - f.R, (f.p - 1) // 2, (f.p + 1) // 2, f.R2]
    vals = edge + [rnd.randrange(f.p) for _ in range(200)]
    a = [x for x in vals for _ in vals[:12]]
    b = [y for _ in vals for y in vals[:12]]
    A, B = f.encode(a), f.encode(b)
    assert f.decode(C.fp_op(fid, "mul", A, B)) == [x * y % f.p for x, y in zip(a, b)]
    assert f.decode(C.fp_op(fid, "add", A, B)) == [(x + y) % f.p for x, y in zip(a, b)]
    assert f.decode(C.fp_op(fid, "sub", A, B)) == [(x - y) % f.p for x, y in zip(a, b)]
    assert f.decode(C.fp_op(fid, "sqr", A)) == [x * x % f.p for x in a]
    assert f.decode(C.fp_op(fid, "dbl", A)) == [2 * x % f.p for x in a]
    assert f.decode(C.fp_op(fid, "neg", A)) == [(-x) % f.p for x in a]
    # into_bigint: Montgomery limbs -> canonical limbs; limb-exact vs the Python REDC restatement
    ib = C.fp_op(fid, "into_bigint", A)
    assert [f.from_limbs(r) for r in ib] == [x % f.p for x in a]
    assert [f.into_bigint_redc(f.to_mont(x)) for x in vals] == vals
    # limb-level CIOS restatement == math
    for x, y in zip(a[:300], b[:300]):
        assert f.mont_mul_cios(f.to_mont(x), f.to_mont(y)) == f.to_mont(x * y % f.p)
    # every result is fully reduced
    assert all(f.from_limbs(r) < f.p for r in C.fp_op(fid, "mul", A, B))


def test_g1_table_pins_group_law(g1_table):
    cv = O.BLS12_381
    acc = None
    for i in range(1000):
        assert g1_table[i] == acc, i
        acc = cv.add(acc, cv.G)
    # XYZZ madd chain (Python restatement) reproduces the table, incl. inf + P and P + P (doubling branch)
    b = cv.xyzz_zero()
    for i in range(1, 200):
        b = cv.xyzz_madd(b, cv.G)
        assert cv.xyzz_to_affine(b) == g1_table[i]
    # C restatement: bucket_i = bucket_{i-1} + G, checked limb-exact after affine conversion
    G = cv.encode_affine([cv.G])
    buck = np.zeros((1, 24), dtype=np.uint64)
    buck[0, :6] = cv.fq.limbs(cv.fq.R)
    buck[0, 6:12] = cv.fq.limbs(cv.fq.R)
    for i in range(1, 300):
        buck = C.ec_op(0, "madd", buck, G)
        aff = C.ec_op(0, "jac_to_affine", C.ec_op(0, "to_jac", buck))
        assert cv.decode_affine(aff)[0] == g1_table[i]


def test_ec_formulas_c_exceptional_cases(g1_table):
    cv = O.BLS12_381
    pts = g1_table
    enc = lambda P: cv.encode_affine([P])

    def bucket_of(P):
        z = np.zeros((1, 24), dtype=np.uint64)
        z[0, :6] = cv.fq.limbs(cv.fq.R)
        z[0, 6:12] = cv.fq.limbs(cv.fq.R)
        return C.ec_op(0, "madd", z, enc(P))

    def aff(b):
        return cv.decode_affine(C.ec_op(0, "jac_to_affine", C.ec_op(0, "to_jac", b)))[0]

    assert aff(C.ec_op(0, "madd", bucket_of(pts[5]), enc(pts[5]))) == pts[10]      # doubling branch
    assert aff(C.ec_op(0, "msub", bucket_of(pts[5]), enc(pts[5]))) is None         # P - P
    assert aff(C.ec_op(0, "madd", bucket_of(pts[5]), enc(None))) == pts[5]         # + infinity
    assert aff(C.ec_op(0, "msub", bucket_of(pts[9]), enc(pts[4]))) == pts[5]
    assert aff(C.ec_op(0, "add", bucket_of(pts[7]), bucket_of(pts[8]))) == pts[15]
    assert aff(C.ec_op(0, "add", bucket_of(pts[7]), bucket_of(pts[7]))) == pts[14]
    assert aff(C.ec_op(0, "add", bucket_of(pts[7]), bucket_of(None))) == pts[7]
    assert aff(C.ec_op(0, "add", bucket_of(None), bucket_of(pts[7]))) == pts[7]
    assert aff(C.ec_op(0, "add", bucket_of(pts[7]), bucket_of(cv.neg(pts[7])))) is None
    assert aff(C.ec_op(0, "dbl", bucket_of(pts[21]))) == pts[42]
    j7 = C.ec_op(0, "to_jac", bucket_of(pts[7]))
    j9 = C.ec_op(0, "to_jac", C.ec_op(0, "madd", bucket_of(pts[4]), enc(pts[5])))  # non-trivial z
    assert cv.decode_affine(C.ec_op(0, "jac_to_affine", C.ec_op(0, "jac_add", j7, j9)))[0] == pts[16]
    assert cv.decode_affine(C.ec_op(0, "jac_to_affine", C.ec_op(0, "jac_dbl", j9)))[0] == pts[18]
    assert cv.decode_affine(C.ec_op(0, "jac_to_affine", C.ec_op(0, "jac_add", j9, j9)))[0] == pts[18]


def test_make_digits_c_vs_python():
    rnd = random.Random(7)
    fr = O.BLS12_381_FR
    for w in (3, 7, 8, 13, 15, 16, 19, 21, 22):
        for _ in range(40):
            s = rnd.randrange(fr.p)
            d = O.make_digits(s, w, 255)
            assert d == C.make_digits(np.array(fr.limbs(s), dtype=np.uint64), w, 255)
            assert sum(x << (i * w) for i, x in enumerate(d)) == s
            assert all(-(1 << (w - 1)) <= x < (1 << (w - 1)) for x in d[:-1]) and d[-1] >= 0
    assert [C.window_size(n) for n in (1, 31, 32, 1 << 10, 1 << 16, 1 << 20, 1 << 24, 1 << 26)] == \
           [O.ark_window(n) for n in (1, 31, 32, 1 << 10, 1 << 16, 1 << 20, 1 << 24, 1 << 26)]
    assert [O.ark_window(1 << k) for k in (16, 20, 24, 26)] == [13, 15, 18, 19]


def test_msm_kat_from_table(g1_table):
    """MSM KATs by linearity of the i*G table: bases {i*G}, small scalars with sum s_i*i < 1000."""
    cv, fr = O.BLS12_381, O.BLS12_381_FR
    idx = [1, 2, 3, 5, 8, 13, 21, 34, 55, 0, 7, 7]
    sc = [3, 1, 4, 1, 5, 9, 2, 6, 5, 11, 0, 2]
    total = sum(i * s for i, s in zip(idx, sc))
    assert total < 1000
    bases = cv.encode_affine([g1_table[i] for i in idx])
    scal = fr.encode(sc)
    for threads in (1, 4):
        assert cv.decode_affine(C.msm_affine(0, bases, scal, threads))[0] == g1_table[total]
    assert cv.decode_affine(C.msm_naive(0, bases, scal))[0] == g1_table[total]
    assert O.pippenger_wnaf(cv, [g1_table[i] for i in idx], sc) == g1_table[total]
    # negative scalars: (r - k) * P = -k * P
    sc2 = [fr.p - 1, 2]
    b2 = cv.encode_affine([g1_table[10], g1_table[30]])
    assert cv.decode_affine(C.msm_affine(0, b2, fr.encode(sc2)))[0] == g1_table[50]


@pytest.mark.parametrize("cid", [0, 1])
def test_msm_random_c_vs_naive_vs_python(cid):
    """test_var_base_msm (test-templates/src/msm.rs:17-32) on the oracles themselves."""
    cv = O.CURVES[cid]
    rnd = random.Random(99 + cid)
    n = 96
    ks = [rnd.randrange(1, cv.fr.p) for _ in range(n)]
    pts = [cv.mul(cv.G, k) for k in ks]
    pts[5] = None                      # identity base
    pts[9] = pts[8]                    # repeated base
    sc = [rnd.randrange(cv.fr.p) for _ in range(n)]
    sc[3] = 0
    sc[4] = 1
    sc[6] = cv.fr.p - 1
    ks[5] = 0
    ks[9] = ks[8]
    want = cv.mul(cv.G, sum(k * s for k, s in zip(ks, sc)) % cv.fr.p)
    assert O.naive_msm(cv, pts, sc) == want
    assert O.pippenger_wnaf(cv, pts, sc, c=4) == want
    B, S = cv.encode_affine(pts), cv.fr.encode(sc)
    assert cv.decode_affine(C.msm_naive(cid, B, S))[0] == want
    for threads, c in ((1, 0), (2, 0), (8, 0), (1, 5), (3, 9)):
        assert cv.decode_affine(C.msm_affine(cid, B, S, threads, c))[0] == want


@pytest.mark.parametrize("fid", [1, 3])
def test_fft_oracles(fid):
    """test_fft_correctness / parallel_fft_consistency (poly/src/domain/radix2/mod.rs:351-391,438-535)."""
    fr = {1: O.BLS12_381_FR, 3: O.BN254_FR}[fid]
    rnd = random.Random(5 + fid)
    for log_n in range(0, 8):
        n = 1 << log_n
        coeffs = [rnd.randrange(fr.p) for _ in range(n)]
        for offset in (1, fr.generator):
            dom = O.Radix2Domain(fr, n, offset)
            want = dom.fft_horner(coeffs)
            assert dom.fft(coeffs) == want
            off = None if offset == 1 else fr.encode([offset])
            got = C.fft(fid, fr.encode(coeffs), False, off, threads=1)
            assert fr.decode(got) == want
            back = C.fft(fid, got, True, off, threads=3)
            assert fr.decode(back) == coeffs
            assert dom.ifft(want) == coeffs
    # larger: C (multi-threaded, exercises root compaction at num_chunks >= 128) vs Python restatement
    n = 1 << 11
    coeffs = [rnd.randrange(fr.p) for _ in range(n)]
    dom = O.Radix2Domain(fr, n)
    want = dom.fft(coeffs)
    assert fr.decode(C.fft(fid, fr.encode(coeffs), False, None, threads=4)) == want
    assert fr.decode(C.fft(fid, fr.encode(want), True, None, threads=4)) == coeffs
    g, gi, ni = C.domain_params(fid, 11)
    assert fr.decode(g)[0] == dom.group_gen and fr.decode(gi)[0] == dom.group_gen_inv and fr.decode(ni)[0] == dom.size_inv


def test_root_of_unity_constants():
    fr = O.BLS12_381_FR
    assert fr.two_adicity == 32
    assert hex(fr.two_adic_root) == "0x16a2a19edfe81f20d09b681922c813b4b63683508c2280b93829971f439f0d2b"
    d = O.Radix2Domain(fr, 1 << 24)
    assert hex(d.group_gen) == "0x291cf6d68823e6876e0bcd91ee76273072cf6a8029b7d7bc92cf4deb77bd779c"
    assert pow(d.group_gen, 1 << 23, fr.p) == fr.p - 1
    assert O.BN254_FR.two_adicity == 28


def test_g2_oracle_pinned_on_reference_table(golden_dir, kats):
    """The G2 oracle (Fq2 arithmetic + Jacobian group law, oracle/pyoracle.py) against the reference's own
    [0*G2 .. 999*G2] table (curves/bls12_381/src/curves/tests/g2_uncompressed_valid_test_vectors.dat) and Fq2 KATs."""
    g2 = O.BLS12_381_G2
    tab = np.load(os.path.join(golden_dir, "bls12_381_g2_multiples.npy"))

    def row(i):
        v = [sum(int(tab[i, 6 * k + j]) << (64 * j) for j in range(6)) for k in range(4)]
        return None if not any(v) else ((v[0], v[1]), (v[2], v[3]))

    assert row(0) is None and row(1) == g2.G
    acc = None
    for i in range(1000):                                  # i*G2 by repeated addition == table
        assert row(i) == acc, i
        acc = g2.add(acc, g2.G)
    rnd = random.Random(9)
    for i in [2, 3, 999] + [rnd.randrange(1000) for _ in range(10)]:
        assert g2.mul(g2.G, i) == row(i)                   # double-and-add
    assert g2.mul(g2.G, g2.fr.p - 5) == g2.neg(row(5)) and g2.mul(g2.G, g2.fr.p) is None
    assert g2.naive_msm([row(3), row(10), None, row(7)], [5, 2, 9, g2.fr.p - 1]) == row(28)
    enc = g2.encode_affine([row(5), None])
    assert g2.decode_affine(enc) == [row(5), None] and not enc[1].any()
    f = kats["bls12_381_fq_field"]
    F = g2.F
    a = tuple(int(x, 16) for x in f["fq2_mul"]["a"]); b = tuple(int(x, 16) for x in f["fq2_mul"]["b"])
    assert F.mul(a, b) == tuple(int(x, 16) for x in f["fq2_mul"]["r"])
    a = tuple(int(x, 16) for x in f["fq2_square"]["a"])
    assert F.sqr(a) == tuple(int(x, 16) for x in f["fq2_square"]["r"])
    a = tuple(int(x, 16) for x in f["fq2_inverse"]["a"])
    assert F.inv(a) == tuple(int(x, 16) for x in f["fq2_inverse"]["r"])
