"""Host-side mirror of the two immediate callers of the NTT in ark-poly:

  DensePolynomial * DensePolynomial   poly/src/polynomial/univariate/dense.rs:641-656  (fft, fft, pointwise, ifft)
  Evaluations::interpolate            poly/src/evaluations/univariate/mod.rs:41-50     (ifft)

Coefficient vectors are (len, 4) uint64 Montgomery limb arrays (numpy: host path; torch CUDA tensors: device-resident —
the three vectors never leave HBM between the steps)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .radix2 import Radix2EvaluationDomain, _is_torch


def _trim(coeffs):
    """DensePolynomial::from_coefficients_vec: drop leading (high-degree) zero coefficients — same rule for host arrays and
    device tensors (the degree is found on the device; only one integer comes back)"""
    if _is_torch(coeffs):
        import torch
        nz = torch.nonzero(coeffs.reshape(-1, 4).ne(0).any(dim=1))
        return coeffs[: (int(nz[-1]) + 1) if nz.numel() else 0]
    nz = np.flatnonzero(coeffs.any(axis=1))
    return coeffs[: (nz[-1] + 1) if nz.size else 0]


def poly_mul(field_id: int, a, b):
    """&DensePolynomial * &DensePolynomial; zero polynomial (empty vector) in -> zero polynomial out."""
    la = (a.numel() if _is_torch(a) else np.asarray(a).size) // 4
    lb = (b.numel() if _is_torch(b) else np.asarray(b).size) // 4
    if la == 0 or lb == 0:   # zero polynomial: an empty vector in the caller's container
        if _is_torch(a):
            import torch
            return torch.zeros((0, 4), dtype=a.dtype, device=a.device)
        return np.zeros((0, 4), dtype=np.uint64)
    n = _lib.lib().b200_poly_mul_size(field_id, la, lb)
    if n == 0:
        raise ValueError("field is not smooth enough to construct domain")   # the reference's expect()
    if _is_torch(a):
        import torch
        out = torch.empty((n, 4), dtype=a.dtype, device=a.device)
        with torch.cuda.device(a.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().b200_poly_mul_fr_dev(field_id, a.data_ptr(), la, b.data_ptr(), lb, out.data_ptr(), st))
        return _trim(out)
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    out = np.empty((n, 4), dtype=np.uint64)
    _lib.check(_lib.lib().b200_poly_mul_fr(field_id, a.ctypes.data_as(ctypes.c_void_p), la, b.ctypes.data_as(ctypes.c_void_p), lb,
                                           out.ctypes.data_as(ctypes.c_void_p)))
    return _trim(out)


def interpolate(domain: Radix2EvaluationDomain, evals):
    """Evaluations::interpolate: ifft over the evaluations' domain, then from_coefficients_vec"""
    return _trim(domain.ifft(evals))
