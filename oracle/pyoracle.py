"""Python big-int oracle (O1) for the arkworks-rs/algebra v0.6.0 hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in `algebra_b200/` may import this module;
only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg do.

This is a *restatement* (math level, Python integers) of the reference
algorithms, each function citing the file:line under /root/reference that it
follows.  It is pinned against the reference's own golden vectors by
`tests/test_oracle_golden.py` (i*G table, Montgomery constants, Fq2 KATs).

Conventions: field elements are canonical Python ints in [0, p) unless a name
says `mont`; limb arrays are little-endian u64 (numpy uint64), Montgomery form
with R = 2^(64*N), exactly the in-memory layout of `ark_ff::Fp`
(ff/src/fields/models/fp/mod.rs:107-115, ff/src/biginteger/mod.rs:34).
"""
from __future__ import annotations

import numpy as np

MASK64 = (1 << 64) - 1


# --------------------------------------------------------------------------
# Field parameters (curves/bls12_381/src/fields/{fq,fr}.rs,
# curves/bn254/src/fields/{fq,fr}.rs).  R, R2, INV are *derived* here the way
# ff/src/fields/models/fp/montgomery_backend.rs:21-24,520-538 derives them.
# --------------------------------------------------------------------------
class Field:
    def __init__(self, name: str, p: int, nlimbs: int, generator: int):
        self.name = name
        self.p = p
        self.N = nlimbs
        self.bits = p.bit_length()
        self.R = (1 << (64 * nlimbs)) % p
        self.R2 = (self.R * self.R) % p
        self.Rinv = pow(self.R, -1, p)
        # INV = -p^{-1} mod 2^64  (montgomery_backend.rs:520-538)
        self.INV = (-pow(p, -1, 1 << 64)) & MASK64
        self.generator = generator
        # two-adicity (ff-macros/src/montgomery/mod.rs:44-55)
        s, t = 0, p - 1
        while t % 2 == 0:
            t //= 2
            s += 1
        self.two_adicity = s
        self.trace = t
        self.two_adic_root = pow(generator, t, p)

    # -- representation ----------------------------------------------------
    def to_mont(self, a: int) -> int:
        return (a * self.R) % self.p

    def from_mont(self, a: int) -> int:
        return (a * self.Rinv) % self.p

    def limbs(self, a: int) -> list[int]:
        return [(a >> (64 * i)) & MASK64 for i in range(self.N)]

    def from_limbs(self, l) -> int:
        return sum(int(x) << (64 * i) for i, x in enumerate(l))

    # arrays: canonical ints -> (n, N) uint64 Montgomery limbs and back
    def encode(self, vals) -> np.ndarray:
        out = np.zeros((len(vals), self.N), dtype=np.uint64)
        for i, v in enumerate(vals):
            m = self.to_mont(v % self.p)
            for j in range(self.N):
                out[i, j] = (m >> (64 * j)) & MASK64
        return out

    def decode(self, arr: np.ndarray) -> list[int]:
        arr = np.asarray(arr, dtype=np.uint64).reshape(-1, self.N)
        return [self.from_mont(self.from_limbs(row)) for row in arr]

    # -- limb-level CIOS restatement (montgomery_backend.rs:214-233) --------
    def mont_mul_cios(self, a_m: int, b_m: int) -> int:
        """a~ * b~ * R^-1 mod p computed limb by limb exactly as the reference's
        no-carry CIOS loop does (mac_with_carry / mac_discard), to pin the
        limb-level algorithm — not just the mathematical result."""
        N, p, INV = self.N, self.limbs(self.p), self.INV
        a, b = self.limbs(a_m), self.limbs(b_m)
        r = [0] * N
        for i in range(N):
            t = r[0] + a[0] * b[i]
            r0, c1 = t & MASK64, t >> 64
            k = (r0 * INV) & MASK64
            c2 = (r0 + k * p[0]) >> 64
            for j in range(1, N):
                t = r[j] + a[j] * b[i] + c1
                rj, c1 = t & MASK64, t >> 64
                t = rj + k * p[j] + c2
                r[j - 1], c2 = t & MASK64, t >> 64
            r[N - 1] = c1 + c2
            assert r[N - 1] <= MASK64  # the "no-carry" claim (34-41)
        v = self.from_limbs(r)
        if v >= self.p:  # __subtract_modulus
            v -= self.p
        return v

    def into_bigint_redc(self, a_m: int) -> int:
        """Montgomery -> canonical by N reduction rounds (montgomery_backend.rs:392-412)."""
        N, p, INV = self.N, self.limbs(self.p), self.INV
        r = self.limbs(a_m)
        for i in range(N):
            k = (r[i] * INV) & MASK64
            carry = (r[i] + k * p[0]) >> 64
            for j in range(1, N):
                t = r[(j + i) % N] + k * p[j] + carry
                r[(j + i) % N], carry = t & MASK64, t >> 64
            r[i % N] = carry
        return self.from_limbs(r)


BLS12_381_FQ = Field(
    "bls12_381_fq",
    0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
    6, 2)
BLS12_381_FR = Field(
    "bls12_381_fr",
    52435875175126190479447740508185965837690552500527637822603658699938581184513, 4, 7)
BN254_FQ = Field(
    "bn254_fq",
    21888242871839275222246405745257275088696311157297823662689037894645226208583, 4, 3)
BN254_FR = Field(
    "bn254_fr",
    21888242871839275222246405745257275088548364400416034343698204186575808495617, 4, 5)


# --------------------------------------------------------------------------
# Curves (curves/bls12_381/src/curves/g1.rs:41-49,199-205;
#         curves/bn254/src/curves/g1.rs:27-42,92-96).  a = 0 for both.
# --------------------------------------------------------------------------
class Curve:
    def __init__(self, cid: int, name: str, fq: Field, fr: Field, b: int, gx: int, gy: int):
        self.id, self.name, self.fq, self.fr, self.b = cid, name, fq, fr, b
        self.G = (gx, gy)
        self.scalar_bits = fr.bits  # MODULUS_BIT_SIZE, variable_base/mod.rs:451
        assert self.on_curve(self.G)

    def on_curve(self, P) -> bool:
        if P is None:
            return True
        x, y = P
        p = self.fq.p
        return (y * y - x * x * x - self.b) % p == 0

    # -- affine group law on canonical ints; None = infinity -----------------
    def neg(self, P):
        return None if P is None else (P[0], (-P[1]) % self.fq.p)

    def add(self, P, Q):
        p = self.fq.p
        if P is None:
            return Q
        if Q is None:
            return P
        x1, y1 = P
        x2, y2 = Q
        if x1 == x2:
            if (y1 + y2) % p == 0:
                return None
            lam = (3 * x1 * x1) * pow(2 * y1, -1, p) % p
        else:
            lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
        x3 = (lam * lam - x1 - x2) % p
        return (x3, (lam * (x1 - x3) - y1) % p)

    def mul(self, P, k: int):
        """double-and-add (ec/src/scalar_mul/mod.rs:27-40) on affine ints."""
        k %= self.fr.p
        # Jacobian internally for speed, same group element
        return self.jac_to_affine(self._jac_mul(P, k))

    # Jacobian helpers (fast path for the naive MSM oracle)
    def _jac_dbl(self, P):
        if P is None:
            return None
        p = self.fq.p
        X, Y, Z = P
        if Y == 0:
            return None
        A = X * X % p
        B = Y * Y % p
        C = B * B % p
        D = 4 * X * B % p
        E = 3 * A % p
        X3 = (E * E - 2 * D) % p
        Y3 = (E * (D - X3) - 8 * C) % p
        Z3 = 2 * Y * Z % p
        return (X3, Y3, Z3)

    def _jac_add(self, P, Q):
        if P is None:
            return Q
        if Q is None:
            return P
        p = self.fq.p
        X1, Y1, Z1 = P
        X2, Y2, Z2 = Q
        Z1Z1 = Z1 * Z1 % p
        Z2Z2 = Z2 * Z2 % p
        U1 = X1 * Z2Z2 % p
        U2 = X2 * Z1Z1 % p
        S1 = Y1 * Z2 * Z2Z2 % p
        S2 = Y2 * Z1 * Z1Z1 % p
        if U1 == U2:
            if S1 == S2:
                return self._jac_dbl(P)
            return None
        H = (U2 - U1) % p
        Rr = (S2 - S1) % p
        HH = H * H % p
        HHH = H * HH % p
        V = U1 * HH % p
        X3 = (Rr * Rr - HHH - 2 * V) % p
        Y3 = (Rr * (V - X3) - S1 * HHH) % p
        Z3 = Z1 * Z2 * H % p
        return (X3, Y3, Z3)

    def _jac_mul(self, P, k):
        if P is None or k == 0:
            return None
        acc = None
        base = (P[0], P[1], 1)
        for bit in bin(k)[2:]:
            acc = self._jac_dbl(acc)
            if bit == "1":
                acc = self._jac_add(acc, base)
        return acc

    def jac_to_affine(self, J):
        """ec/src/models/short_weierstrass/affine.rs:374-396"""
        if J is None or J[2] == 0:
            return None
        p = self.fq.p
        X, Y, Z = J
        zi = pow(Z, -1, p)
        zi2 = zi * zi % p
        return (X * zi2 % p, Y * zi2 * zi % p)

    # -- limb encodings ------------------------------------------------------
    def encode_affine(self, pts) -> np.ndarray:
        """list of affine points -> (n, 2N) u64 Montgomery limbs; infinity = (0,0)
        (affine.rs:91-104, ZeroFlag=(): short_weierstrass/mod.rs:224-230)."""
        N = self.fq.N
        out = np.zeros((len(pts), 2 * N), dtype=np.uint64)
        for i, P in enumerate(pts):
            if P is None:
                continue
            out[i, :N] = self.fq.limbs(self.fq.to_mont(P[0]))
            out[i, N:] = self.fq.limbs(self.fq.to_mont(P[1]))
        return out

    def decode_affine(self, arr):
        N = self.fq.N
        arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 2 * N)
        pts = []
        for row in arr:
            xm = self.fq.from_limbs(row[:N])
            ym = self.fq.from_limbs(row[N:])
            if xm == 0 and ym == 0:
                pts.append(None)
            else:
                pts.append((self.fq.from_mont(xm), self.fq.from_mont(ym)))
        return pts

    def decode_jacobian(self, arr):
        """3N u64 limbs (x,y,z Montgomery; z==0 => infinity, group.rs:145-151) -> affine."""
        N = self.fq.N
        arr = np.asarray(arr, dtype=np.uint64).reshape(3 * N)
        x, y, z = (self.fq.from_mont(self.fq.from_limbs(arr[i * N:(i + 1) * N])) for i in range(3))
        return self.jac_to_affine((x, y, z))

    # -- XYZZ bucket formulas, restated (bucket.rs) ----------------------------
    # bucket = (X, Y, ZZ, ZZZ) canonical ints; zero = (1,1,0,0) (bucket.rs:78-83)
    def xyzz_zero(self):
        return (1, 1, 0, 0)

    def xyzz_is_zero(self, B):
        return B[2] == 0 and B[3] == 0  # bucket.rs:108-110

    def xyzz_mdbl(self, P):
        """affine.rs:169-201 (mdbl-2008-s-1)"""
        if P is None:
            return self.xyzz_zero()
        p = self.fq.p
        x, y = P
        U = 2 * y % p
        V = U * U % p
        W = U * V % p
        S = x * V % p
        M = 3 * x * x % p
        X3 = (M * M - 2 * S) % p
        Y3 = (M * (S - X3) - W * y) % p
        return (X3, Y3, V, W)

    def xyzz_madd(self, B, P):
        """Bucket += Affine, bucket.rs:168-238 (madd-2008-s) incl. exceptional cases."""
        if P is None:
            return B
        if self.xyzz_is_zero(B):
            return (P[0], P[1], 1, 1)
        p = self.fq.p
        X1, Y1, ZZ1, ZZZ1 = B
        x2, y2 = P
        U2 = x2 * ZZ1 % p
        S2 = y2 * ZZZ1 % p
        if X1 == U2:
            if Y1 == S2:
                return self.xyzz_mdbl(P)
            return self.xyzz_zero()
        Pp = (U2 - X1) % p
        Rr = (S2 - Y1) % p
        PP = Pp * Pp % p
        PPP = Pp * PP % p
        Q = X1 * PP % p
        X3 = (Rr * Rr - PPP - 2 * Q) % p
        Y3 = (Rr * (Q - X3) - Y1 * PPP) % p
        return (X3, Y3, ZZ1 * PP % p, ZZZ1 * PPP % p)

    def xyzz_dbl(self, B):
        """bucket.rs:112-146 (dbl-2008-s-1), a = 0"""
        p = self.fq.p
        X, Y, ZZ, ZZZ = B
        U = 2 * Y % p
        V = U * U % p
        W = U * V % p
        S = X * V % p
        M = 3 * X * X % p
        X3 = (M * M - 2 * S) % p
        Y3 = (M * (S - X3) - W * Y) % p
        return (X3, Y3, V * ZZ % p, W * ZZZ % p)

    def xyzz_add(self, A, B):
        """Bucket += &Bucket, bucket.rs:256-337 (add-2008-s)."""
        if self.xyzz_is_zero(A):
            return B
        if self.xyzz_is_zero(B):
            return A
        p = self.fq.p
        X1, Y1, ZZ1, ZZZ1 = A
        X2, Y2, ZZ2, ZZZ2 = B
        U1 = X1 * ZZ2 % p
        U2 = X2 * ZZ1 % p
        S1 = Y1 * ZZZ2 % p
        S2 = Y2 * ZZZ1 % p
        if U1 == U2:
            if S1 == S2:
                return self.xyzz_dbl(A)
            return self.xyzz_zero()
        Pp = (U2 - U1) % p
        Rr = (S2 - S1) % p
        PP = Pp * Pp % p
        PPP = Pp * PP % p
        Q = U1 * PP % p
        X3 = (Rr * Rr - PPP - 2 * Q) % p
        Y3 = (Rr * (Q - X3) - S1 * PPP) % p
        return (X3, Y3, ZZ1 * ZZ2 * PP % p, ZZZ1 * ZZZ2 * PPP % p)

    def xyzz_to_jac(self, B):
        """From<Bucket> for Projective, bucket.rs:389-398"""
        if self.xyzz_is_zero(B):
            return None
        p = self.fq.p
        X, Y, ZZ, ZZZ = B
        return (X * ZZ % p, Y * ZZZ % p, ZZ)

    def xyzz_to_affine(self, B):
        return self.jac_to_affine(self.xyzz_to_jac(B))


BLS12_381 = Curve(
    0, "bls12_381_g1", BLS12_381_FQ, BLS12_381_FR, 4,
    3685416753713387016781088315183077757961620795782546409894578378688607592378376318836054947676345821548104185464507,
    1339506544944476473020471379941921221584933875938349620426543736416511423956333506472724655353366534992391756441569)
BN254 = Curve(1, "bn254_g1", BN254_FQ, BN254_FR, 3, 1, 2)
CURVES = {0: BLS12_381, 1: BN254}
FR_FIELDS = {0: BLS12_381_FR, 1: BN254_FR}


# --------------------------------------------------------------------------
# MSM
# --------------------------------------------------------------------------
def naive_msm(curve: Curve, bases, scalars):
    """Σ s_i·P_i by double-and-add — test-templates/src/msm.rs:8-15."""
    acc = None
    for P, s in zip(bases, scalars):
        if P is None or s % curve.fr.p == 0:
            continue
        acc = curve._jac_add(acc, curve._jac_mul(P, s % curve.fr.p))
    return curve.jac_to_affine(acc)


def log2_ceil(n: int) -> int:
    """ark_std::log2: ceil(log2(n)), 0 for n<=1."""
    if n <= 1:
        return 0
    return (n - 1).bit_length()


def ln_without_floats(a: int) -> int:
    """ec/src/scalar_mul/mod.rs:22-25"""
    return log2_ceil(a) * 69 // 100


def ark_window(size: int) -> int:
    """variable_base/mod.rs:445-449"""
    return 3 if size < 32 else ln_without_floats(size) + 2


def make_digits(scalar: int, w: int, num_bits: int, nlimbs: int = 4) -> list[int]:
    """variable_base/mod.rs:754-794, limb-exact (incl. the single/two-limb read rule)."""
    limbs = [(scalar >> (64 * i)) & MASK64 for i in range(nlimbs)]
    radix = 1 << w
    window_mask = radix - 1
    carry = 0
    if num_bits == 0:
        num_bits = scalar.bit_length()
    digits_count = -(-num_bits // w)
    out = []
    for i in range(digits_count):
        bit_offset = i * w
        u64_idx = bit_offset // 64
        bit_idx = bit_offset % 64
        if bit_idx < 64 - w or u64_idx == nlimbs - 1:
            bit_buf = limbs[u64_idx] >> bit_idx
        else:
            bit_buf = ((limbs[u64_idx] >> bit_idx) | (limbs[1 + u64_idx] << (64 - bit_idx))) & MASK64
        coef = carry + (bit_buf & window_mask)
        carry = (coef + radix // 2) >> w
        digit = coef - (carry << w)
        if i == digits_count - 1:
            digit += carry << w
        out.append(digit)
    return out


def pippenger_wnaf(curve: Curve, bases, scalars, c: int | None = None):
    """msm_bigint_wnaf_parallel restated (variable_base/mod.rs:437-503) with XYZZ
    buckets; returns an affine point."""
    size = min(len(bases), len(scalars))
    if size == 0:
        return None
    if c is None:
        c = ark_window(size)
    num_bits = curve.scalar_bits
    digits_count = -(-num_bits // c)
    digs = [make_digits(s % curve.fr.p, c, num_bits) for s in scalars[:size]]
    window_sums = []
    for i in range(digits_count):
        buckets = [curve.xyzz_zero() for _ in range(1 << c)]
        for d, P in zip(digs, bases):
            s = d[i]
            if s > 0:
                buckets[s - 1] = curve.xyzz_madd(buckets[s - 1], P)
            elif s < 0:
                buckets[-s - 1] = curve.xyzz_madd(buckets[-s - 1], curve.neg(P))
        running = curve.xyzz_zero()
        res = curve.xyzz_zero()
        for b in reversed(buckets):
            running = curve.xyzz_add(running, b)
            res = curve.xyzz_add(res, running)
        window_sums.append(res)
    lowest = curve.xyzz_to_jac(window_sums[0])
    total = None
    for s in reversed(window_sums[1:]):
        total = curve._jac_add(total, curve.xyzz_to_jac(s))
        for _ in range(c):
            total = curve._jac_dbl(total)
    return curve.jac_to_affine(curve._jac_add(lowest, total))


# --------------------------------------------------------------------------
# NTT (poly/src/domain/radix2)
# --------------------------------------------------------------------------
class Radix2Domain:
    """Radix2EvaluationDomain::new / get_coset, poly/src/domain/radix2/mod.rs:55-92."""

    def __init__(self, fr: Field, num_coeffs: int, offset: int = 1):
        size = 1 if num_coeffs <= 1 else 1 << (num_coeffs - 1).bit_length()
        log = size.bit_length() - 1
        if log > fr.two_adicity:
            raise ValueError("domain too large")
        self.fr, self.size, self.log_size = fr, size, log
        p = fr.p
        # get_root_of_unity, ff/src/fields/fft_friendly.rs:66-82
        g = fr.two_adic_root
        for _ in range(log, fr.two_adicity):
            g = g * g % p
        self.group_gen = g
        self.group_gen_inv = pow(g, -1, p)
        self.size_inv = pow(size % p, -1, p)
        self.offset = offset % p
        self.offset_inv = pow(offset, -1, p)
        self.offset_pow_size = pow(offset, size, p)

    def element(self, i: int) -> int:
        return self.offset * pow(self.group_gen, i, self.fr.p) % self.fr.p

    # math-level definition (radix2/mod.rs:370-372): out[i] = poly(element(i))
    def fft_horner(self, coeffs):
        p = self.fr.p
        out = []
        for i in range(self.size):
            x = self.element(i)
            acc = 0
            for cf in reversed(coeffs):
                acc = (acc * x + cf) % p
            out.append(acc)
        return out

    # restatement of the reference's algorithm (fft.rs:74-79,90-102,252-295,373-380)
    def fft(self, coeffs):
        p, n = self.fr.p, self.size
        x = list(coeffs[:n]) + [0] * (n - min(len(coeffs), n))  # resize, radix2/mod.rs:144
        if self.offset != 1:  # distribute_powers, domain/mod.rs:115-128
            pw = 1
            for i in range(n):
                x[i] = x[i] * pw % p
                pw = pw * self.offset % p
        self._io_helper(x, self.group_gen)
        _derange(x, self.log_size)
        return x

    def ifft(self, evals):
        p, n = self.fr.p, self.size
        x = list(evals[:n]) + [0] * (n - min(len(evals), n))
        _derange(x, self.log_size)
        self._oi_helper(x, self.group_gen_inv)
        if self.offset == 1:  # fft.rs:83-87
            return [v * self.size_inv % p for v in x]
        pw = self.size_inv
        for i in range(n):
            x[i] = x[i] * pw % p
            pw = pw * self.offset_inv % p
        return x

    def _roots(self, root):
        p = self.fr.p
        r, out = 1, []
        for _ in range(self.size // 2):
            out.append(r)
            r = r * root % p
        return out

    def _io_helper(self, x, root):
        """DIF: lo' = lo+hi ; hi' = (lo-hi)*w  (fft.rs:190-198,252-295)"""
        p, n = self.fr.p, len(x)
        roots = self._roots(root)
        gap = n // 2
        while gap > 0:
            nchunks = n // (2 * gap)
            for c0 in range(0, n, 2 * gap):
                for j in range(gap):
                    lo, hi = x[c0 + j], x[c0 + j + gap]
                    x[c0 + j] = (lo + hi) % p
                    x[c0 + j + gap] = (lo - hi) * roots[j * nchunks] % p
            gap //= 2

    def _oi_helper(self, x, root):
        """DIT: hi*=w ; lo'=lo+hi ; hi'=lo-hi  (fft.rs:201-210,297-349)"""
        p, n = self.fr.p, len(x)
        roots = self._roots(root)
        gap = 1
        while gap < n:
            nchunks = n // (2 * gap)
            for c0 in range(0, n, 2 * gap):
                for j in range(gap):
                    hi = x[c0 + j + gap] * roots[j * nchunks] % p
                    lo = x[c0 + j]
                    x[c0 + j] = (lo + hi) % p
                    x[c0 + j + gap] = (lo - hi) % p
            gap *= 2


def _bitrev(a: int, log_len: int) -> int:
    return int(format(a, "0%db" % log_len)[::-1], 2) if log_len else 0


def _derange(x, log_len):
    """fft.rs:373-380"""
    for idx in range(1, len(x) - 1):
        r = _bitrev(idx, log_len)
        if idx < r:
            x[idx], x[r] = x[r], x[idx]


# --------------------------------------------------------------------------
# deterministic input generation shared by tests / bench (own PRNG: splitmix64)
# --------------------------------------------------------------------------
def splitmix64(seed: int):
    s = seed & MASK64
    while True:
        s = (s + 0x9E3779B97F4A7C15) & MASK64
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        yield z ^ (z >> 31)


def rand_field(field: Field, n: int, seed: int) -> list[int]:
    """uniform in [0,p): 64N random bits masked to MODULUS_BIT_SIZE, rejection
    (ff/src/fields/models/fp/mod.rs:521-548)."""
    g = splitmix64(seed)
    mask = (1 << field.bits) - 1
    out = []
    while len(out) < n:
        v = 0
        for i in range(field.N):
            v |= next(g) << (64 * i)
        v &= mask
        if v < field.p:
            out.append(v)
    return out


# --------------------------------------------------------------------------
# G2 of BLS12-381 (curves/bls12_381/src/curves/g2.rs:54-90,218-236): y^2 = x^3 + 4(u + 1) over Fq2 = Fq[u]/(u^2 + 1)
# (curves/bls12_381/src/fields/fq2.rs:10-24).  Elements are pairs (c0, c1) of canonical ints; the limb encoding is
# QuadExtField's in-memory order c0 | c1 (ff/src/fields/models/quadratic_extension.rs:105-113), each 6 Montgomery u64.
# Checker-only code: math-level restatement (Jacobian over Fq2), pinned on the reference's own i*G2 table
# (tests/test_oracle_golden.py).
# --------------------------------------------------------------------------
class Fq2:
    def __init__(self, fq: Field):
        self.fq, self.p = fq, fq.p

    def add(self, a, b):
        return ((a[0] + b[0]) % self.p, (a[1] + b[1]) % self.p)

    def sub(self, a, b):
        return ((a[0] - b[0]) % self.p, (a[1] - b[1]) % self.p)

    def neg(self, a):
        return ((-a[0]) % self.p, (-a[1]) % self.p)

    def mul(self, a, b):
        p = self.p
        return ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)

    def sqr(self, a):
        return self.mul(a, a)

    def smul(self, a, k: int):
        return (a[0] * k % self.p, a[1] * k % self.p)

    def inv(self, a):
        p = self.p
        n = pow((a[0] * a[0] + a[1] * a[1]) % p, -1, p)
        return (a[0] * n % p, (-a[1]) * n % p)

    def is_zero(self, a):
        return a[0] % self.p == 0 and a[1] % self.p == 0


class CurveG2:
    def __init__(self, cid: int, name: str, fq: Field, fr: Field, b, gx, gy):
        self.id, self.name, self.fq, self.fr, self.b = cid, name, fq, fr, b
        self.F = Fq2(fq)
        self.G = (gx, gy)
        self.scalar_bits = fr.bits
        assert self.on_curve(self.G)

    def on_curve(self, P) -> bool:
        if P is None:
            return True
        F = self.F
        x, y = P
        return F.is_zero(F.sub(F.sqr(y), F.add(F.mul(F.sqr(x), x), self.b)))

    def neg(self, P):
        return None if P is None else (P[0], self.F.neg(P[1]))

    def _jac_dbl(self, P):
        if P is None:
            return None
        F = self.F
        X, Y, Z = P
        if F.is_zero(Y):
            return None
        A, B = F.sqr(X), F.sqr(Y)
        C = F.sqr(B)
        D = F.smul(F.mul(X, B), 4)
        E = F.smul(A, 3)
        X3 = F.sub(F.sqr(E), F.smul(D, 2))
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), F.smul(C, 8))
        return (X3, Y3, F.smul(F.mul(Y, Z), 2))

    def _jac_add(self, P, Q):
        if P is None:
            return Q
        if Q is None:
            return P
        F = self.F
        X1, Y1, Z1 = P
        X2, Y2, Z2 = Q
        Z1Z1, Z2Z2 = F.sqr(Z1), F.sqr(Z2)
        U1, U2 = F.mul(X1, Z2Z2), F.mul(X2, Z1Z1)
        S1, S2 = F.mul(F.mul(Y1, Z2), Z2Z2), F.mul(F.mul(Y2, Z1), Z1Z1)
        if U1 == U2:
            return self._jac_dbl(P) if S1 == S2 else None
        H, Rr = F.sub(U2, U1), F.sub(S2, S1)
        HH = F.sqr(H)
        HHH = F.mul(H, HH)
        V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.sqr(Rr), HHH), F.smul(V, 2))
        Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
        return (X3, Y3, F.mul(F.mul(Z1, Z2), H))

    def jac_to_affine(self, J):
        if J is None or self.F.is_zero(J[2]):
            return None
        F = self.F
        zi = F.inv(J[2])
        zi2 = F.sqr(zi)
        return (F.mul(J[0], zi2), F.mul(F.mul(J[1], zi2), zi))

    def add(self, P, Q):
        lift = lambda A: None if A is None else (A[0], A[1], (1, 0))
        return self.jac_to_affine(self._jac_add(lift(P), lift(Q)))

    def mul(self, P, k: int):
        k %= self.fr.p
        if P is None or k == 0:
            return None
        acc, base = None, (P[0], P[1], (1, 0))
        for bit in bin(k)[2:]:
            acc = self._jac_dbl(acc)
            if bit == "1":
                acc = self._jac_add(acc, base)
        return self.jac_to_affine(acc)

    # -- limb encodings: affine = x.c0 | x.c1 | y.c0 | y.c1 (24 u64), infinity = all zero --------------------------------
    def encode_affine(self, pts) -> np.ndarray:
        N = self.fq.N
        out = np.zeros((len(pts), 4 * N), dtype=np.uint64)
        for i, P in enumerate(pts):
            if P is None:
                continue
            for k, v in enumerate((P[0][0], P[0][1], P[1][0], P[1][1])):
                out[i, k * N:(k + 1) * N] = self.fq.limbs(self.fq.to_mont(v))
        return out

    def decode_affine(self, arr):
        N = self.fq.N
        arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 4 * N)
        pts = []
        for row in arr:
            v = [self.fq.from_limbs(row[k * N:(k + 1) * N]) for k in range(4)]
            if not any(v):
                pts.append(None)
            else:
                c = [self.fq.from_mont(t) for t in v]
                pts.append(((c[0], c[1]), (c[2], c[3])))
        return pts

    def decode_jacobian(self, arr):
        N = self.fq.N
        arr = np.asarray(arr, dtype=np.uint64).reshape(6 * N)
        v = [self.fq.from_mont(self.fq.from_limbs(arr[k * N:(k + 1) * N])) for k in range(6)]
        return self.jac_to_affine(((v[0], v[1]), (v[2], v[3]), (v[4], v[5])))

    def naive_msm(self, bases, scalars):
        acc = None
        for P, s in zip(bases, scalars):
            Q = self.mul(P, s)
            if Q is not None:
                acc = self._jac_add(acc, (Q[0], Q[1], (1, 0)))
        return self.jac_to_affine(acc)


BLS12_381_G2 = CurveG2(
    2, "bls12_381_g2", BLS12_381_FQ, BLS12_381_FR, (4, 4),
    (352701069587466618187139116011060144890029952792775240219908644239793785735715026873347600343865175952761926303160,
     3059144344244213709971259814753781636986470325476647558659373206291635324768958432433509563104347017837885763365758),
    (1985150602287291935568054521177171638300868978215655730859378665066344726373823718423869104263333984641494340347905,
     927553665492332455747201965776037880757740193453592970025027978793976877002675564980949289727957565575433344219582))
CURVES[2] = BLS12_381_G2
