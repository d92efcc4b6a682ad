"""Host-side mirror of the producers of MSM bases:

  batch_mul(curve, base, scalars)      ScalarMul::batch_mul / BatchMulPreprocessing (ec/src/scalar_mul/mod.rs:53-245):
                                       [s_0*B, s_1*B, ...] as affine points
  normalize_batch(curve, projective)   Projective::normalize_batch (ec/src/models/short_weierstrass/group.rs:302-319)

numpy in -> numpy out (staged through device memory), torch CUDA tensors in -> torch CUDA tensor out."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .params import CURVES, G1Curve


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _to_dev(x):
    import torch
    if _is_torch(x):
        return x, True
    a = np.ascontiguousarray(x, dtype=np.uint64)
    return torch.from_numpy(a.view(np.int64)).cuda(), False


def batch_mul(curve: G1Curve | int, base, scalars):
    import torch
    cv = CURVES[curve] if isinstance(curve, int) else curve
    base = np.ascontiguousarray(base, dtype=np.uint64).reshape(2 * cv.N)
    d_s, was_torch = _to_dev(scalars)
    n = d_s.numel() // 4
    out = torch.empty((n, 2 * cv.N), dtype=torch.int64, device=d_s.device)
    with torch.cuda.device(d_s.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().b200_g1_batch_mul_dev(cv.cid, base.ctypes.data_as(ctypes.c_void_p), d_s.data_ptr(), n, out.data_ptr(), st))
    return out if was_torch else out.cpu().numpy().view(np.uint64)


def normalize_batch(curve: G1Curve | int, points_xyz):
    import torch
    cv = CURVES[curve] if isinstance(curve, int) else curve
    d_p, was_torch = _to_dev(points_xyz)
    n = d_p.numel() // (3 * cv.N)
    out = torch.empty((n, 2 * cv.N), dtype=torch.int64, device=d_p.device)
    with torch.cuda.device(d_p.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().b200_g1_normalize_batch_dev(cv.cid, d_p.data_ptr(), n, out.data_ptr(), st))
    return out if was_torch else out.cpu().numpy().view(np.uint64)
