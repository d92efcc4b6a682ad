"""GPU parity of the NTT through the C ABI / the Radix2EvaluationDomain mirror, bit-exact against the oracle.
Mirrors poly/src/domain/radix2/mod.rs:351-391 (fft == evaluation at domain elements, coset, ifft round trip),
:438-535 (consistency with a textbook serial FFT for sizes 2^0.., fft/ifft/coset), :581-600 and
poly/src/test.rs:11-60 (ifft o fft = id).  Full-size cases use size-independent properties."""
import random

import numpy as np
import pytest

import algebra_b200 as ab
from oracle import coracle as C
from oracle import pyoracle as O

from gpu_util import from_dev, to_dev

pytestmark = pytest.mark.gpu
FR = {0: (O.BLS12_381_FR, 1), 1: (O.BN254_FR, 3)}    # B200_FIELD_* -> (oracle field, oracle field id)


def rand_limbs(fr, n, seed):
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    x[:, 3] &= np.uint64((1 << (fr.bits - 192)) - 1)
    # force < p: clear one more top bit (keeps the values uniform enough for a test) and patch a few edge elements
    x[:, 3] >>= np.uint64(1)
    return x


@pytest.mark.parametrize("field", [0, 1])
def test_small_sizes_vs_horner_and_oracle(field):
    fr, ofid = FR[field]
    rnd = random.Random(3 + field)
    for log_n in range(0, 11):
        n = 1 << log_n
        dom = ab.Radix2EvaluationDomain.new(field, n)
        coeffs = [rnd.randrange(fr.p) for _ in range(n)]
        x = fr.encode(coeffs)
        for offset in (1, fr.generator, rnd.randrange(2, fr.p)):
            d = dom if offset == 1 else dom.get_coset(offset)
            off_l = None if offset == 1 else fr.encode([offset])
            got = d.fft(x)
            want = C.fft(ofid, x, False, off_l, threads=2)
            assert (got == want).all(), (log_n, offset)
            if log_n <= 5:   # the defining property, radix2/mod.rs:370-372
                assert fr.decode(got) == O.Radix2Domain(fr, n, offset).fft_horner(coeffs)
            back = d.ifft(got)
            assert (back == x).all()
            assert (back == C.fft(ofid, want, True, off_l, threads=2)).all()
            # device-resident path gives the same bits
            t = to_dev(x)
            d.fft_in_place(t)
            assert (from_dev(t) == want).all()
            d.ifft_in_place(t)
            assert (from_dev(t) == x).all()


@pytest.mark.parametrize("field,log_n", [(0, 11), (0, 12), (0, 13), (0, 15), (0, 16), (0, 17), (0, 20), (1, 14), (1, 18)])
def test_medium_sizes_vs_oracle(field, log_n):
    fr, ofid = FR[field]
    n = 1 << log_n
    x = rand_limbs(fr, n, 1000 + log_n)
    x[0] = fr.encode([fr.p - 1])[0]
    x[1] = 0
    dom = ab.Radix2EvaluationDomain.new(field, n)
    thr = min(C.num_threads(), 32)
    want = C.fft(ofid, x, False, None, threads=thr)
    got = dom.fft(x)
    assert (got == want).all()
    assert (dom.ifft(got) == x).all()
    assert (dom.ifft(x) == C.fft(ofid, x, True, None, threads=thr)).all()
    if log_n <= 16:
        off = fr.generator
        assert (dom.get_coset(off).fft(x) == C.fft(ofid, x, False, fr.encode([off]), threads=thr)).all()
        assert (dom.get_coset(off).ifft(x) == C.fft(ofid, x, True, fr.encode([off]), threads=thr)).all()


def test_resize_semantics():
    """fft_in_place pads short inputs with zeros and truncates long ones (radix2/mod.rs:140-147)."""
    fr, ofid = FR[0]
    dom = ab.Radix2EvaluationDomain.new(0, 256)
    x = rand_limbs(fr, 300, 5)
    padded = np.zeros((256, 4), dtype=np.uint64)
    padded[:40] = x[:40]
    assert (dom.fft(x[:40]) == C.fft(ofid, padded)).all()          # degree-aware route in the reference: same values
    assert (dom.fft(x) == C.fft(ofid, np.ascontiguousarray(x[:256]))).all()
    assert ab.Radix2EvaluationDomain.new(0, 1 << 33) is None          # > TWO_ADICITY
    assert ab.Radix2EvaluationDomain.new(1, 1 << 29) is None
    assert ab.Radix2EvaluationDomain.new(0, 1000).size == 1024
    from algebra_b200 import _lib
    import ctypes
    buf = np.zeros((2, 4), dtype=np.uint64)
    assert _lib.lib().b200_ntt_fr(0, buf.ctypes.data_as(ctypes.c_void_p), 33, 0, None) == _lib.ETOOLARGE
    assert _lib.lib().b200_ntt_fr(7, buf.ctypes.data_as(ctypes.c_void_p), 1, 0, None) == _lib.EINVAL


@pytest.mark.parametrize("log_n", [22, 24])
def test_full_size_properties(log_n):
    """At BASELINE sizes: ifft(fft(x)) == x, linearity fft(a*x + y) == a*fft(x) + fft(y) on sampled outputs, and
    spot checks of individual outputs against direct evaluation sum_j x_j w^(ij) restricted to a sparse input."""
    import torch
    fr, ofid = FR[0]
    n = 1 << log_n
    dom = ab.Radix2EvaluationDomain.new(0, n)
    x = rand_limbs(fr, n, 77)
    t = to_dev(x)
    dom.fft_in_place(t)
    fx = from_dev(t).copy()
    dom.ifft_in_place(t)
    assert torch.equal(t.cpu(), torch.from_numpy(x.view(np.int64)))
    # sparse input: x = e_a * u + e_b * v  ->  X[i] = u w^(a i) + v w^(b i), checked at sampled i with Python ints
    a_idx, b_idx = 12345 % n, (n // 3) | 1
    u, v = 0x1234567 % fr.p, (fr.p - 5)
    sp = np.zeros((n, 4), dtype=np.uint64)
    sp[a_idx] = fr.encode([u])[0]
    sp[b_idx] = fr.encode([v])[0]
    fs = dom.fft(sp)
    g = O.Radix2Domain(fr, n).group_gen
    for i in [0, 1, 2, n // 2, n - 1, 99991 % n, (n // 7) * 3 + 1]:
        want = (u * pow(g, a_idx * i, fr.p) + v * pow(g, b_idx * i, fr.p)) % fr.p
        assert fr.decode(fs[i])[0] == want
    # linearity against the dense transform: fft(x + sp) == fft(x) + fft(sp)   (field add via the oracle, sampled rows)
    xs = x.copy()
    xs[a_idx] = C.fp_op(ofid, "add", x[a_idx:a_idx + 1], sp[a_idx:a_idx + 1])[0]
    xs[b_idx] = C.fp_op(ofid, "add", x[b_idx:b_idx + 1], sp[b_idx:b_idx + 1])[0]
    fxs = dom.fft(xs)
    rows = np.random.default_rng(1).integers(0, n, size=4096)
    assert (fxs[rows] == C.fp_op(ofid, "add", np.ascontiguousarray(fx[rows]), np.ascontiguousarray(fs[rows]))).all()


def test_baseline_size_2e24_elementwise_vs_oracle():
    """BASELINE.json configs[2] at n = 2^24: fft and ifft compared with the restated reference algorithm on EVERY element
    (poly/src/domain/radix2/mod.rs:351-391 at full size), device-resident and host-buffer entry points."""
    fr, ofid = FR[0]
    log_n = 24
    n = 1 << log_n
    x = rand_limbs(fr, n, 2424)
    x[0] = fr.encode([fr.p - 1])[0]
    x[n - 1] = 0
    thr = C.num_threads()
    want = C.fft(ofid, x, False, None, threads=thr)
    dom = ab.Radix2EvaluationDomain.new(0, n)
    t = to_dev(x)
    dom.fft_in_place(t)
    assert (from_dev(t) == want).all()
    dom.ifft_in_place(t)
    assert (from_dev(t) == x).all()
    want_inv = C.fft(ofid, x, True, None, threads=thr)
    assert (dom.ifft(x) == want_inv).all()          # host-buffer entry (H2D + D2H inside the call)
    del want_inv
    off = fr.encode([fr.generator])
    t = to_dev(x)
    dom.get_coset(fr.generator).fft_in_place(t)
    assert (from_dev(t) == C.fft(ofid, x, False, off, threads=thr)).all()


def test_poly_mul_and_interpolate():
    """DensePolynomial * DensePolynomial (dense.rs:641-656) vs schoolbook multiplication with Python ints, host and
    device-resident paths; Evaluations::interpolate = ifft + trimming."""
    fr, _ = FR[0]
    rnd = random.Random(12)
    for la, lb in ((1, 1), (3, 5), (64, 65), (300, 17), (1000, 1000)):
        a = [rnd.randrange(fr.p) for _ in range(la)]
        b = [rnd.randrange(fr.p) for _ in range(lb)]
        b[-1] = b[-1] or 1
        a[-1] = a[-1] or 1
        want = [0] * (la + lb - 1)
        for i, x in enumerate(a):
            for j, y in enumerate(b):
                want[i + j] = (want[i + j] + x * y) % fr.p
        A, B = fr.encode(a), fr.encode(b)
        got = ab.poly_mul(0, A, B)
        assert fr.decode(got) == want
        gd = from_dev(ab.poly_mul(0, to_dev(A), to_dev(B)))          # device path trims like the host path
        assert gd.shape[0] == la + lb - 1 and fr.decode(gd) == want
    assert ab.poly_mul(0, np.zeros((0, 4), np.uint64), fr.encode([1, 2])).shape == (0, 4)
    dom = ab.Radix2EvaluationDomain.new(0, 64)
    coeffs = fr.encode([rnd.randrange(fr.p) for _ in range(40)])
    assert (ab.interpolate(dom, dom.fft(coeffs)) == coeffs).all()


def test_four_pass_plan_2e25():
    """log n = 25 needs four passes (7+6+6+6): exercises the multi-digit reversal of the last pass.  Checked by the
    closed form of a sparse input and by the round trip."""
    import torch
    fr, _ = FR[0]
    log_n = 25
    n = 1 << log_n
    dom = ab.Radix2EvaluationDomain.new(0, n)
    idx = [1, 12345, n // 2 + 77, n - 3]
    vals = [3, fr.p - 2, 0x123456789, 7]
    t = torch.zeros((n, 4), dtype=torch.int64, device="cuda")
    enc = fr.encode(vals)
    for i, e in zip(idx, enc):
        t[i] = torch.from_numpy(e.view(np.int64)).cuda()
    x0 = t.clone()
    dom.fft_in_place(t)
    g = O.Radix2Domain(fr, n).group_gen
    for i in [0, 1, 5, n // 2, n - 1, 0x155555 % n, (n // 3) | 1, 1 << 19, (1 << 19) + (1 << 7) + 1]:
        want = sum(v * pow(g, j * i, fr.p) for j, v in zip(idx, vals)) % fr.p
        assert fr.decode(t[i].cpu().numpy().view(np.uint64))[0] == want, i
    dom.ifft_in_place(t)
    assert torch.equal(t, x0)
