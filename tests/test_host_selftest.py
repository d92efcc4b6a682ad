"""CPU check of the *device* arithmetic headers (algebra_b200/csrc/fp.cuh, ec.cuh): tools/host_selftest.cpp
compiles them with g++ through the PTX-emulation path and this test compares every op, limb-exact, with the
C oracle (which is pinned on the reference's golden vectors).  The same comparisons run on the real GPU in
tests/test_gpu_primitives.py."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import coracle as C
from oracle import pyoracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(ROOT, "tools", "libhost_selftest.so")
    src = os.path.join(ROOT, "tools", "host_selftest.cpp")
    deps = [src] + [os.path.join(ROOT, "algebra_b200", "csrc", f) for f in ("fp.cuh", "ec.cuh", "field_consts.inc")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-o", so, src])
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))


FIELDS = [(0, O.BLS12_381_FQ), (1, O.BLS12_381_FR), (2, O.BN254_FQ), (3, O.BN254_FR)]


@pytest.mark.parametrize("fid,f", FIELDS)
def test_fp_ops(lib, fid, f):
    rnd = random.Random(fid)
    edge = [0, 1, 2, f.p - 1, f.p - 2, f.R, f.p - f.R, (f.p - 1) // 2, f.R2]
    vals = edge + [rnd.randrange(f.p) for _ in range(200)]
    a = [x for x in vals for _ in vals[:10]]
    b = [y for _ in vals for y in vals[:10]]
    A, B = f.encode(a), f.encode(b)
    for op in ("mul", "add", "sub", "sqr", "dbl", "neg", "into_bigint"):
        want = C.fp_op(fid, op, A, B)
        out = np.empty_like(A)
        assert lib.selftest_fp_op(fid, C.OPS[op], _p(A.view(np.uint32)), _p(B.view(np.uint32)), _p(out.view(np.uint32)), len(a)) == 0
        assert (out == want).all(), op
    out = np.empty_like(A)   # symmetric squaring variant (op 9) == mul(a, a)
    assert lib.selftest_fp_op(fid, 9, _p(A.view(np.uint32)), _p(A.view(np.uint32)), _p(out.view(np.uint32)), len(a)) == 0
    assert (out == C.fp_op(fid, "sqr", A)).all()
    if fid in (0, 2):        # rolled-row multiplication variant (BlsFqRolled / BnFqRolled) == the unrolled one
        out = np.empty_like(A)
        assert lib.selftest_fp_op({0: 4, 2: 5}[fid], 0, _p(A.view(np.uint32)), _p(B.view(np.uint32)), _p(out.view(np.uint32)), len(a)) == 0
        assert (out == C.fp_op(fid, "mul", A, B)).all()
    nz = np.ascontiguousarray(A[1:40])   # windowed / symmetric-squaring inversion (block-shared inversion of the MSM) == a^(p-2)
    out = np.empty_like(nz)
    assert lib.selftest_fp_op(fid, 10, _p(nz.view(np.uint32)), _p(nz.view(np.uint32)), _p(out.view(np.uint32)), len(nz)) == 0
    assert (out == C.fp_op(fid, "inv", nz)).all()
    canon = np.ascontiguousarray(C.fp_op(fid, "into_bigint", A))
    out = np.empty_like(A)
    lib.selftest_fp_op(fid, 7, _p(canon.view(np.uint32)), _p(canon.view(np.uint32)), _p(out.view(np.uint32)), len(a))
    assert (out == A).all()
    A3 = np.ascontiguousarray(A[10:60])
    out = np.empty_like(A3)
    lib.selftest_fp_op(fid, 8, _p(A3.view(np.uint32)), _p(A3.view(np.uint32)), _p(out.view(np.uint32)), len(A3))
    assert (out == C.fp_op(fid, "inv", A3)).all()


@pytest.mark.parametrize("cid", [0, 1])
def test_ec_ops(lib, cid):
    cv = O.CURVES[cid]
    N = cv.fq.N
    rnd = random.Random(10 + cid)
    ks = [rnd.randrange(1, 1 << 40) for _ in range(24)]
    pts = [cv.mul(cv.G, k) for k in ks]
    aff = cv.encode_affine(pts + [None])
    zero = np.zeros((1, 4 * N), dtype=np.uint64)
    zero[0, :N] = cv.fq.limbs(cv.fq.R)
    zero[0, N:2 * N] = cv.fq.limbs(cv.fq.R)

    def run(op, a, b=None):
        code, wa, wb, wo = C.EC_OPS[op]
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, wa * N)
        bb = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, wb * N) if wb else a
        out = np.empty((a.shape[0], wo * N), dtype=np.uint64)
        assert lib.selftest_ec_op(cid, code, _p(a.view(np.uint32)), _p(bb.view(np.uint32)), _p(out.view(np.uint32)), a.shape[0]) == 0
        want = C.ec_op(cid, op, a, bb if wb else None)
        assert (out == want).all(), op
        return out

    # buckets with non-trivial zz/zzz: acc_i = P_i + P_{i+1} (+ P_{i+2})
    n = len(pts)
    b1 = run("madd", np.repeat(zero, n, 0), aff[:n])                # inf + P   (copy branch)
    b2 = run("madd", b1, np.roll(aff[:n], 1, 0))                     # generic
    b3 = run("msub", b2, np.roll(aff[:n], 2, 0))                     # generic, negated
    run("madd", b1, aff[:n])                                         # P + P     (doubling branch)
    run("msub", b1, aff[:n])                                         # P - P     (-> infinity)
    run("madd", b3, np.repeat(aff[n:n + 1], n, 0))                   # + infinity
    run("add", b2, b3)
    run("add", b2, b2)                                               # doubling branch
    run("add", b3, np.repeat(zero, n, 0))
    run("add", np.repeat(zero, n, 0), b3)
    neg = b3.copy().reshape(n, 4, N)
    neg[:, 1, :] = C.fp_op({0: 0, 1: 2}[cid], "neg", np.ascontiguousarray(neg[:, 1, :]))
    run("add", b3, neg.reshape(n, -1))                               # -> infinity
    run("dbl", b3)
    j2, j3 = run("to_jac", b2), run("to_jac", b3)
    run("to_jac", np.repeat(zero, 2, 0))
    run("jac_add", j2, j3)
    run("jac_add", j2, j2)
    run("jac_dbl", j3)
    a3 = run("jac_to_affine", j3)
    want = [cv.add(cv.add(pts[i], pts[(i - 1) % n]), cv.neg(pts[(i - 2) % n])) for i in range(n)]
    assert cv.decode_affine(a3) == want


def test_g2_ec_ops_vs_oracle(lib):
    """ec.cuh instantiated over Fp2 (the G2 MSM's point arithmetic): XYZZ mixed add / add / double, Jacobian add / double and the
    conversions, run on the host through the PTX emulation and compared (as affine points) with the G2 oracle, which is pinned on
    the reference's i*G2 table."""
    g2 = O.BLS12_381_G2
    fq = O.BLS12_381_FQ
    rnd = random.Random(77)
    ks = [rnd.randrange(1, 1 << 40) for _ in range(12)]
    P = [g2.mul(g2.G, k) for k in ks]
    Q = [g2.mul(g2.G, rnd.randrange(1, 1 << 40)) for _ in ks]
    Q[3] = P[3]                      # doubling branch
    Q[4] = g2.neg(P[4])              # cancellation
    Q[5] = None                      # + identity

    def enc2(c):                     # Fq2 -> 12 u64 Montgomery limbs
        return np.array(fq.limbs(fq.to_mont(c[0])) + fq.limbs(fq.to_mont(c[1])), dtype=np.uint64)

    def xyzz(Pt):                    # affine -> bucket (x, y, 1, 1); identity -> (1, 1, 0, 0)
        if Pt is None:
            return np.concatenate([enc2((1, 0)), enc2((1, 0)), enc2((0, 0)), enc2((0, 0))])
        return np.concatenate([enc2(Pt[0]), enc2(Pt[1]), enc2((1, 0)), enc2((1, 0))])

    def run(op, a, b, wo):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64) if b is not None else a
        out = np.zeros((a.shape[0], wo * 12), dtype=np.uint64)
        assert lib.selftest_ec_op(2, op, _p(a.view(np.uint32)), _p(b.view(np.uint32)), _p(out.view(np.uint32)), a.shape[0]) == 0
        return out

    def xyzz_affine(rows):           # bucket -> jacobian (op 4) -> affine (op 5)
        return g2.decode_affine(run(5, run(4, rows, None, 3), None, 2))

    B = np.stack([xyzz(p) for p in P])
    Aq = g2.encode_affine(Q)
    assert xyzz_affine(run(0, B, Aq, 4)) == [g2.add(p, q) for p, q in zip(P, Q)]              # Bucket += Affine
    assert xyzz_affine(run(1, B, Aq, 4)) == [g2.add(p, g2.neg(q)) for p, q in zip(P, Q)]      # Bucket -= Affine
    B2 = run(0, np.stack([xyzz(None)] * len(Q)), Aq, 4)                                        # identity bucket += Q
    assert xyzz_affine(B2) == Q
    S = run(0, B, g2.encode_affine(P[1:] + P[:1]), 4)                                          # non-trivial zz, zzz
    T = run(0, np.stack([xyzz(q) for q in Q]), g2.encode_affine(P[2:] + P[:2]), 4)
    sa, ta = xyzz_affine(S), xyzz_affine(T)
    assert xyzz_affine(run(2, S, T, 4)) == [g2.add(x, y) for x, y in zip(sa, ta)]              # Bucket += &Bucket
    assert xyzz_affine(run(2, S, S, 4)) == [g2.add(x, x) for x in sa]                          # ... doubling branch
    assert xyzz_affine(run(3, S, None, 4)) == [g2.add(x, x) for x in sa]                       # double_in_place
    J, K = run(4, S, None, 3), run(4, T, None, 3)
    assert g2.decode_affine(run(5, run(6, J, K, 3), None, 2)) == [g2.add(x, y) for x, y in zip(sa, ta)]   # Projective += Projective
    assert g2.decode_affine(run(5, run(6, J, J, 3), None, 2)) == [g2.add(x, x) for x in sa]
    assert g2.decode_affine(run(5, run(7, J, None, 3), None, 2)) == [g2.add(x, x) for x in sa]             # double_in_place
