"""GPU parity of the device arithmetic (fp.cuh / ec.cuh) through the C ABI's element-wise kernels, limb-exact against
the C oracle.  Mirrors test-templates/src/fields.rs:143-272 (ring ops on random + edge operands) and
test-templates/src/groups.rs:286-360 (Bucket += Affine / += Bucket vs plain addition)."""
import random

import numpy as np
import pytest

from oracle import coracle as C
from oracle import pyoracle as O

from gpu_util import ec_op, fp_op

pytestmark = pytest.mark.gpu
FIELDS = [(0, O.BLS12_381_FQ), (1, O.BLS12_381_FR), (2, O.BN254_FQ), (3, O.BN254_FR)]


@pytest.mark.parametrize("fid,f", FIELDS)
def test_fp_ops(fid, f):
    rnd = random.Random(100 + fid)
    edge = [0, 1, 2, f.p - 1, f.p - 2, f.R, f.p - f.R, (f.p - 1) // 2, (f.p + 1) // 2, f.R2]
    vals = edge + [rnd.randrange(f.p) for _ in range(500)]
    a = [x for x in vals for _ in vals[:12]]
    b = [y for _ in vals for y in vals[:12]]
    A, B = f.encode(a), f.encode(b)
    for op in ("mul", "add", "sub", "sqr", "dbl", "neg", "into_bigint"):
        got = fp_op(fid, C.OPS[op], A, B)
        assert (got == C.fp_op(fid, op, A, B)).all(), op
    canon = np.ascontiguousarray(C.fp_op(fid, "into_bigint", A))
    assert (fp_op(fid, 7, canon) == A).all()
    A3 = np.ascontiguousarray(A[12:140])
    assert (fp_op(fid, 8, A3) == C.fp_op(fid, "inv", A3)).all()
    # a larger random batch so every SM runs the kernel
    big_a = [rnd.randrange(f.p) for _ in range(1 << 14)]
    big_b = [rnd.randrange(f.p) for _ in range(1 << 14)]
    BA, BB = f.encode(big_a), f.encode(big_b)
    assert (fp_op(fid, 0, BA, BB) == C.fp_op(fid, "mul", BA, BB)).all()
    # repeated squaring chain (micro-benchmark mode) agrees with pow
    got = f.decode(fp_op(fid, 3, BA[:64], None, reps=10))
    assert got == [pow(x, 1 << 10, f.p) for x in big_a[:64]]


@pytest.mark.parametrize("cid", [0, 1])
def test_ec_ops(cid):
    cv = O.CURVES[cid]
    N = cv.fq.N
    rnd = random.Random(10 + cid)
    n = 96
    ks = [rnd.randrange(1, 1 << 60) for _ in range(n)]
    pts = [cv.mul(cv.G, k) for k in ks]
    aff = cv.encode_affine(pts + [None])
    zero = np.zeros((1, 4 * N), dtype=np.uint64)
    zero[0, :N] = cv.fq.limbs(cv.fq.R)
    zero[0, N:2 * N] = cv.fq.limbs(cv.fq.R)

    def run(op, a, b=None):
        code, wa, wb, wo = C.EC_OPS[op]
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, wa * N)
        bb = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, wb * N) if wb else None
        got = ec_op(cid, code, a, bb, wo)
        want = C.ec_op(cid, op, a, bb)
        assert (got == want).all(), op
        return got

    b1 = run("madd", np.repeat(zero, n, 0), aff[:n])
    b2 = run("madd", b1, np.roll(aff[:n], 1, 0))
    b3 = run("msub", b2, np.roll(aff[:n], 2, 0))
    run("madd", b1, aff[:n])                              # doubling branch
    run("msub", b1, aff[:n])                              # -> infinity
    run("madd", b3, np.repeat(aff[n:n + 1], n, 0))        # + infinity
    run("add", b2, b3)
    run("add", b2, b2)
    run("add", b3, np.repeat(zero, n, 0))
    run("add", np.repeat(zero, n, 0), b3)
    neg = b3.copy().reshape(n, 4, N)
    neg[:, 1, :] = C.fp_op({0: 0, 1: 2}[cid], "neg", np.ascontiguousarray(neg[:, 1, :]))
    run("add", b3, neg.reshape(n, -1))
    run("dbl", b3)
    j2, j3 = run("to_jac", b2), run("to_jac", b3)
    run("to_jac", np.repeat(zero, 2, 0))
    run("jac_add", j2, j3)
    run("jac_add", j2, j2)
    run("jac_dbl", j3)
    a3 = run("jac_to_affine", j3)
    want = [cv.add(cv.add(pts[i], pts[(i - 1) % n]), cv.neg(pts[(i - 2) % n])) for i in range(n)]
    assert cv.decode_affine(a3) == want
