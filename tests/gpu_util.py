"""helpers for the -m gpu tests: device buffers come from torch (plumbing only), every compute call goes through
the C ABI of libalgebra_b200.so."""
import ctypes

import numpy as np


def to_dev(arr: np.ndarray):
    import torch
    a = np.ascontiguousarray(arr, dtype=np.uint64)
    return torch.from_numpy(a.view(np.int64)).cuda()


def from_dev(t) -> np.ndarray:
    return t.cpu().numpy().view(np.uint64)


def dev_empty(shape):
    import torch
    return torch.empty(shape, dtype=torch.int64, device="cuda")


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def fp_op(field: int, op: int, a: np.ndarray, b: np.ndarray | None = None, reps: int = 1) -> np.ndarray:
    from algebra_b200 import _lib
    da = to_dev(a)
    db = to_dev(b) if b is not None else None
    out = dev_empty(da.shape)
    n = a.size // {0: 6, 1: 4, 2: 4, 3: 4}[field]
    _lib.check(_lib.lib().b200_fp_op_dev(field, op, da.data_ptr(), db.data_ptr() if db is not None else None,
                                         out.data_ptr(), n, reps, stream()))
    return from_dev(out).reshape(a.shape)


def ec_op(curve: int, op: int, a: np.ndarray, b: np.ndarray | None, out_width: int) -> np.ndarray:
    from algebra_b200 import _lib
    N = {0: 6, 1: 4}[curve]
    da = to_dev(a)
    db = to_dev(b) if b is not None else None
    n = a.shape[0]
    out = dev_empty((n, out_width * N))
    _lib.check(_lib.lib().b200_ec_op_dev(curve, op, da.data_ptr(), db.data_ptr() if db is not None else None,
                                         out.data_ptr(), n, stream()))
    return from_dev(out)
