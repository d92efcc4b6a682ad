"""ctypes binding of oracle/libark_oracle.so (the limb-level C restatement, O2).

TEST INFRASTRUCTURE ONLY — see the header of ark_oracle.c.  Imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libark_oracle.so")

FIELD_BLS_FQ, FIELD_BLS_FR, FIELD_BN_FQ, FIELD_BN_FR = 0, 1, 2, 3
FIELD_LIMBS = {0: 6, 1: 4, 2: 4, 3: 4}
CURVE_LIMBS = {0: 6, 1: 4}
OPS = {"mul": 0, "add": 1, "sub": 2, "sqr": 3, "dbl": 4, "neg": 5, "into_bigint": 6, "from_bigint": 7, "inv": 8}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ark_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        u64p = ctypes.POINTER(ctypes.c_uint64)
        _lib.ark_fp_op.argtypes = [ctypes.c_int, ctypes.c_int, u64p, u64p, u64p, ctypes.c_size_t]
        _lib.ark_ec_op.argtypes = [ctypes.c_int, ctypes.c_int, u64p, u64p, u64p, ctypes.c_size_t]
        _lib.ark_msm.argtypes = [ctypes.c_int, u64p, u64p, ctypes.c_size_t, u64p, ctypes.c_int, ctypes.c_int]
        _lib.ark_msm_naive.argtypes = [ctypes.c_int, u64p, u64p, ctypes.c_size_t, u64p]
        _lib.ark_fft.argtypes = [ctypes.c_int, u64p, ctypes.c_uint, ctypes.c_int, u64p, ctypes.c_int]
        _lib.ark_domain_params.argtypes = [ctypes.c_int, ctypes.c_uint, u64p, u64p, u64p]
        _lib.ark_make_digits.argtypes = [u64p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64)]
        _lib.ark_window_size.argtypes = [ctypes.c_size_t]
        _lib.ark_field_constants.argtypes = [ctypes.c_int, u64p, u64p, u64p, u64p]
    return _lib


def _p(a: np.ndarray | None):
    if a is None:
        return None
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))


def num_threads() -> int:
    return lib().ark_num_threads()


def fp_op(field: int, op: str, a: np.ndarray, b: np.ndarray | None = None) -> np.ndarray:
    N = FIELD_LIMBS[field]
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, N)
    if b is not None:
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, N)
    out = np.empty_like(a)
    rc = lib().ark_fp_op(field, OPS[op], _p(a), _p(b if b is not None else a), _p(out), a.shape[0])
    assert rc == 0
    return out


EC_OPS = {"madd": (0, 4, 2, 4), "msub": (1, 4, 2, 4), "add": (2, 4, 4, 4), "dbl": (3, 4, 0, 4),
          "to_jac": (4, 4, 0, 3), "jac_to_affine": (5, 3, 0, 2), "jac_add": (6, 3, 3, 3), "jac_dbl": (7, 3, 0, 3)}


def ec_op(curve: int, op: str, a: np.ndarray, b: np.ndarray | None = None) -> np.ndarray:
    N = CURVE_LIMBS[curve]
    code, wa, wb, wo = EC_OPS[op]
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, wa * N)
    if wb:
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, wb * N)
    out = np.empty((a.shape[0], wo * N), dtype=np.uint64)
    rc = lib().ark_ec_op(curve, code, _p(a), _p(b) if wb else None, _p(out), a.shape[0])
    assert rc == 0
    return out


def msm(curve: int, bases: np.ndarray, scalars: np.ndarray, threads: int = 1, c: int = 0) -> np.ndarray:
    """reference-algorithm MSM; returns 3N Jacobian Montgomery limbs."""
    N = CURVE_LIMBS[curve]
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 2 * N)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    n = min(bases.shape[0], scalars.shape[0])
    out = np.empty(3 * N, dtype=np.uint64)
    rc = lib().ark_msm(curve, _p(bases), _p(scalars), n, _p(out), threads, c)
    assert rc == 0
    return out


def msm_affine(curve: int, bases, scalars, threads: int = 1, c: int = 0) -> np.ndarray:
    """MSM then into_affine: 2N Montgomery limbs ((0,0) = infinity) — the parity format."""
    j = msm(curve, bases, scalars, threads, c)
    return ec_op(curve, "jac_to_affine", j.reshape(1, -1)).reshape(-1)


def msm_naive(curve: int, bases, scalars) -> np.ndarray:
    N = CURVE_LIMBS[curve]
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 2 * N)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    out = np.empty(2 * N, dtype=np.uint64)
    rc = lib().ark_msm_naive(curve, _p(bases), _p(scalars), min(len(bases), len(scalars)), _p(out))
    assert rc == 0
    return out


def fft(field: int, data: np.ndarray, inverse: bool = False, offset: np.ndarray | None = None,
        threads: int = 1) -> np.ndarray:
    """Radix2EvaluationDomain fft/ifft on n = 2^k Montgomery Fr elements; returns a new array."""
    x = np.array(data, dtype=np.uint64, copy=True).reshape(-1, 4)
    n = x.shape[0]
    assert n & (n - 1) == 0 and n > 0
    if offset is not None:
        offset = np.ascontiguousarray(offset, dtype=np.uint64).reshape(4)
    rc = lib().ark_fft(field, _p(x), n.bit_length() - 1, int(inverse), _p(offset), threads)
    assert rc == 0, rc
    return x


def domain_params(field: int, log_n: int):
    g, gi, ni = (np.empty(4, dtype=np.uint64) for _ in range(3))
    rc = lib().ark_domain_params(field, log_n, _p(g), _p(gi), _p(ni))
    assert rc == 0
    return g, gi, ni


def make_digits(scalar_limbs, w: int, num_bits: int) -> list[int]:
    s = np.ascontiguousarray(scalar_limbs, dtype=np.uint64).reshape(4)
    out = (ctypes.c_int64 * 256)()
    cnt = lib().ark_make_digits(_p(s), w, num_bits, out)
    return [out[i] for i in range(cnt)]


def window_size(n: int) -> int:
    return lib().ark_window_size(n)


def field_constants(field: int):
    N = FIELD_LIMBS[field]
    p, R, R2 = (np.empty(N, dtype=np.uint64) for _ in range(3))
    inv = np.empty(1, dtype=np.uint64)
    assert lib().ark_field_constants(field, _p(p), _p(R), _p(R2), _p(inv)) == 0
    return p, R, R2, int(inv[0])
