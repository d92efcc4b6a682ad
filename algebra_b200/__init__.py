"""algebra_b200 — B200-native (sm_100a) backend for the arkworks-rs/algebra data-parallel hot path:
variable-base MSM over short-Weierstrass G1 and the radix-2 NTT, behind the C ABI of include/algebra_b200.h.
This package is only the host-side mirror of the reference interfaces; all arithmetic runs in
libalgebra_b200.so (hand-written CUDA).  No CPU fallback."""
from . import _lib, params                                    # noqa: F401
from .radix2 import Radix2EvaluationDomain                    # noqa: F401
from .polynomial import interpolate, poly_mul                  # noqa: F401
from .fixed_base import batch_mul, normalize_batch             # noqa: F401
from .variable_base import (ChunkedPippenger, HashMapPippenger, LengthMismatch, MsmStream, bases_free, bases_upload, device_count,   # noqa: F401
                            into_affine, msm, msm_bigint, msm_chunks, msm_multi, msm_u1, msm_u8, msm_u16, msm_u32, msm_u64, msm_unchecked,
                            msm_with_bases, set_bucket_slice, sum_points)
from .params import BLS12_381_G1, BLS12_381_G2, BN254_G1      # noqa: F401

__version__ = "0.1.0"
