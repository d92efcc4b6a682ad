"""Host-side mirror of `ark_ec::VariableBaseMSM` for short-Weierstrass G1 (ec/src/scalar_mul/variable_base/mod.rs:37-151)
over the C ABI.  Same names, argument meaning and error behaviour as the reference:

  msm(bases, scalars)            -> Jacobian limbs, or raises LengthMismatch(min_len) like `Err(min_len)` (:73-77)
  msm_unchecked(bases, scalars)  -> silently truncates to the shorter input (:59-64)

bases:   (n, 2N) uint64 Montgomery affine points, numpy (host) or torch CUDA tensor (device-resident, dtype int64/uint64)
scalars: (n, 4)  uint64 Montgomery Fr elements, same container kind as `bases`
result:  numpy (3N,) uint64 = Projective (x, y, z), Montgomery limbs; `into_affine` gives the comparison form."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .params import CURVES, G1Curve


class LengthMismatch(ValueError):
    """`Err(min_len)` of VariableBaseMSM::msm"""

    def __init__(self, min_len: int):
        super().__init__(f"bases and scalars differ in length; min_len = {min_len}")
        self.min_len = min_len


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _rows(x, width: int) -> int:
    if _is_torch(x):
        return x.numel() // width
    return np.asarray(x).size // width


_KIND_DTYPE = {_lib.SCALARS_FR_MONT: (np.uint64, 4), _lib.SCALARS_BIGINT: (np.uint64, 4), _lib.SCALARS_U8: (np.uint8, 1),
               _lib.SCALARS_U16: (np.uint16, 1), _lib.SCALARS_U32: (np.uint32, 1), _lib.SCALARS_U64: (np.uint64, 1)}


def _msm_kind(curve: G1Curve | int, kind: int, bases, scalars) -> np.ndarray:
    cv = CURVES[curve] if isinstance(curve, int) else curve
    N = cv.N
    dtype, width = _KIND_DTYPE[kind]
    n = min(_rows(bases, 2 * N), _rows(scalars, width))
    out = np.zeros(3 * N, dtype=np.uint64)
    outp = out.ctypes.data_as(ctypes.c_void_p)
    if _is_torch(bases) != _is_torch(scalars):
        raise TypeError("bases and scalars must both be numpy arrays or both be torch CUDA tensors")
    if _is_torch(bases):
        import torch
        if not (bases.is_cuda and scalars.is_cuda and bases.is_contiguous() and scalars.is_contiguous()):
            raise TypeError("device path needs contiguous CUDA tensors")
        if bases.element_size() != 8 or scalars.element_size() != np.dtype(dtype).itemsize:
            raise TypeError("device path: bases must be 8-byte integers and scalars must match the scalar kind's width")
        with torch.cuda.device(bases.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().b200_msm_sw_g1_scalars_dev(cv.cid, kind, bases.data_ptr(), scalars.data_ptr(), n, outp, st))
    else:
        b = np.ascontiguousarray(np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 2 * N)[:n])
        s = np.ascontiguousarray(np.ascontiguousarray(scalars, dtype=dtype).reshape(-1, width)[:n])
        _lib.check(_lib.lib().b200_msm_sw_g1_scalars(cv.cid, kind, b.ctypes.data_as(ctypes.c_void_p), s.ctypes.data_as(ctypes.c_void_p), n, outp))
    return out


def msm_unchecked(curve: G1Curve | int, bases, scalars) -> np.ndarray:
    return _msm_kind(curve, _lib.SCALARS_FR_MONT, bases, scalars)


def msm_bigint(curve: G1Curve | int, bases, bigints) -> np.ndarray:
    """VariableBaseMSM::msm_bigint (:80-85): scalars are canonical `BigInt<4>` limbs ((n, 4) uint64, not Montgomery)."""
    return _msm_kind(curve, _lib.SCALARS_BIGINT, bases, bigints)


def msm_u1(curve, bases, scalars) -> np.ndarray:
    """msm_u1 (:89-91): boolean scalars (one byte each, like Rust's `&[bool]`)."""
    s = scalars if _is_torch(scalars) else np.ascontiguousarray(scalars).astype(np.uint8)
    return _msm_kind(curve, _lib.SCALARS_U8, bases, s)


def msm_u8(curve, bases, scalars) -> np.ndarray:
    return _msm_kind(curve, _lib.SCALARS_U8, bases, scalars)


def msm_u16(curve, bases, scalars) -> np.ndarray:
    return _msm_kind(curve, _lib.SCALARS_U16, bases, scalars)


def msm_u32(curve, bases, scalars) -> np.ndarray:
    return _msm_kind(curve, _lib.SCALARS_U32, bases, scalars)


def msm_u64(curve, bases, scalars) -> np.ndarray:
    return _msm_kind(curve, _lib.SCALARS_U64, bases, scalars)


class MsmStream:
    """b200_msm_stream_*: one set of buckets for a stream of chunks.  Every push is copied to the device while the previous
    chunk is still being accumulated; the bucket reduction and the window combine run once, in finish()."""

    def __init__(self, curve: G1Curve | int, max_chunk: int, n_total_hint: int = 0, kind: int = _lib.SCALARS_FR_MONT):
        self.cv = CURVES[curve] if isinstance(curve, int) else curve
        self.kind = kind
        self._h = ctypes.c_void_p()
        _lib.check(_lib.lib().b200_msm_stream_begin(self.cv.cid, kind, n_total_hint, max(1, max_chunk), ctypes.byref(self._h)))

    def push(self, bases, scalars) -> None:
        dtype, width = _KIND_DTYPE[self.kind]
        b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 2 * self.cv.N)
        s = np.ascontiguousarray(scalars, dtype=dtype).reshape(-1, width)
        n = min(len(b), len(s))
        if self._h is None:
            raise RuntimeError("stream already finished")
        _lib.check(_lib.lib().b200_msm_stream_push(self._h, b.ctypes.data_as(ctypes.c_void_p), s.ctypes.data_as(ctypes.c_void_p), n))

    def finish(self) -> np.ndarray:
        out = np.zeros(3 * self.cv.N, dtype=np.uint64)
        h, self._h = self._h, None
        _lib.check(_lib.lib().b200_msm_stream_finish(h, out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def __del__(self):
        if getattr(self, "_h", None) is not None:
            _lib.lib().b200_msm_stream_abort(self._h)
            self._h = None


def msm_chunks(curve: G1Curve | int, bases_stream, scalars_stream, step: int = 1 << 20) -> np.ndarray:
    """VariableBaseMSM::msm_chunks (:119-150): the scalar stream may be shorter than the base stream; the LAST
    len(scalars) bases are used (`skip(bases.len() - scalars.len())`), consumed `step` pairs at a time through the
    streaming C ABI (one bucket set, one reduction at the end — the reference folds a full MSM per chunk)."""
    cv = CURVES[curve] if isinstance(curve, int) else curve
    nb, ns = _rows(bases_stream, 2 * cv.N), _rows(scalars_stream, 4)
    if ns > nb:
        raise ValueError("scalars_stream.len() <= bases_stream.len()")
    b = np.asarray(bases_stream).reshape(-1, 2 * cv.N)[nb - ns:]
    s = np.asarray(scalars_stream).reshape(-1, 4)
    if ns == 0:
        return msm_unchecked(cv, b[:0], s[:0])
    st = MsmStream(cv, min(step, ns), ns)
    for lo in range(0, ns, step):
        st.push(b[lo:lo + step], s[lo:lo + step])
    return st.finish()


class ChunkedPippenger:
    """stream_pippenger.rs:10-66: buffer (base, bigint scalar) pairs, hand every full buffer of `buf_size` pairs to the device
    (b200_msm_stream_push with canonical BigInt scalars) and finalize once."""

    def __init__(self, curve: G1Curve | int, max_msm_buffer: int):
        self.cv = CURVES[curve] if isinstance(curve, int) else curve
        self.buf_size = max_msm_buffer
        self.bases, self.scalars = [], []
        self._stream = None

    def add(self, base, scalar_bigint):
        self.bases.append(np.asarray(base, dtype=np.uint64).reshape(2 * self.cv.N))
        self.scalars.append(np.asarray(scalar_bigint, dtype=np.uint64).reshape(4))
        if len(self.scalars) == self.buf_size:
            self._flush()

    def _flush(self):
        if self._stream is None:
            self._stream = MsmStream(self.cv, self.buf_size, 0, _lib.SCALARS_BIGINT)
        self._stream.push(np.stack(self.bases), np.stack(self.scalars))
        self.bases, self.scalars = [], []

    def finalize(self) -> np.ndarray:
        if self.scalars:
            self._flush()
        if self._stream is None:
            return msm_unchecked(self.cv, np.zeros((0, 2 * self.cv.N), np.uint64), np.zeros((0, 4), np.uint64))
        st, self._stream = self._stream, None
        return st.finish()


def device_count() -> int:
    return _lib.lib().b200_device_count()


def msm_multi(curve: G1Curve | int, bases, scalars, ngpus: int) -> np.ndarray:
    """b200_msm_sw_g1_multi: host arrays in, one process drives `ngpus` devices (input-chunk sharding, partial sums added on
    device 0).  Same truncation rule as msm_unchecked."""
    cv = CURVES[curve] if isinstance(curve, int) else curve
    b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 2 * cv.N)
    s = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    n = min(len(b), len(s))
    out = np.zeros(3 * cv.N, dtype=np.uint64)
    _lib.check(_lib.lib().b200_msm_sw_g1_multi(cv.cid, ngpus, b.ctypes.data_as(ctypes.c_void_p), s.ctypes.data_as(ctypes.c_void_p), n,
                                               out.ctypes.data_as(ctypes.c_void_p)))
    return out


class BasesHandle:
    def __init__(self, cv, h, n):
        self.cv, self.h, self.n = cv, h, n


def bases_upload(curve: G1Curve | int, bases, ngpus: int = 1) -> BasesHandle:
    """b200_bases_upload: keep `bases` resident in HBM (sharded over `ngpus` devices) across MSM calls."""
    cv = CURVES[curve] if isinstance(curve, int) else curve
    b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 2 * cv.N)
    h = ctypes.c_void_p()
    _lib.check(_lib.lib().b200_bases_upload(cv.cid, ngpus, b.ctypes.data_as(ctypes.c_void_p), len(b), ctypes.byref(h)))
    return BasesHandle(cv, h, len(b))


def msm_with_bases(handle: BasesHandle, scalars, kind: int = _lib.SCALARS_FR_MONT) -> np.ndarray:
    dtype, width = _KIND_DTYPE[kind]
    s = np.ascontiguousarray(scalars, dtype=dtype).reshape(-1, width)
    n = min(len(s), handle.n)
    out = np.zeros(3 * handle.cv.N, dtype=np.uint64)
    _lib.check(_lib.lib().b200_msm_bases(handle.h, kind, s.ctypes.data_as(ctypes.c_void_p), n, out.ctypes.data_as(ctypes.c_void_p)))
    return out


def bases_free(handle: BasesHandle) -> None:
    if handle.h is not None:
        _lib.lib().b200_bases_free(handle.h)
        handle.h = None


def msm(curve: G1Curve | int, bases, scalars) -> np.ndarray:
    cv = CURVES[curve] if isinstance(curve, int) else curve
    nb, ns = _rows(bases, 2 * cv.N), _rows(scalars, 4)
    if nb != ns:
        raise LengthMismatch(min(nb, ns))
    return msm_unchecked(cv, bases, scalars)


def into_affine(curve: G1Curve | int, xyz: np.ndarray) -> np.ndarray:
    """Projective -> Affine (ec/src/models/short_weierstrass/affine.rs:374-396); identity -> (0,0)."""
    cv = CURVES[curve] if isinstance(curve, int) else curve
    xyz = np.ascontiguousarray(xyz, dtype=np.uint64).reshape(3 * cv.N)
    out = np.zeros(2 * cv.N, dtype=np.uint64)
    _lib.check(_lib.lib().b200_g1_into_affine(cv.cid, xyz.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p)))
    return out


def sum_points(curve: G1Curve | int, points_xyz: np.ndarray) -> np.ndarray:
    """sum of k Projective points ((k, 3N) uint64) — `Sum for Projective`, used after the multi-GPU gather."""
    cv = CURVES[curve] if isinstance(curve, int) else curve
    pts = np.ascontiguousarray(points_xyz, dtype=np.uint64).reshape(-1, 3 * cv.N)
    out = np.zeros(3 * cv.N, dtype=np.uint64)
    _lib.check(_lib.lib().b200_g1_sum(cv.cid, pts.ctypes.data_as(ctypes.c_void_p), pts.shape[0], out.ctypes.data_as(ctypes.c_void_p)))
    return out


def set_window(c: int) -> None:
    """Pippenger window override (0 = automatic) — for the window sweep of BASELINE configs[1]."""
    _lib.check(_lib.lib().b200_set_msm_window(c))


def set_affine_levels(levels: int) -> None:
    """batched-affine pre-reduction levels (0 = off, -1 = automatic); result-neutral tuning knob"""
    _lib.check(_lib.lib().b200_set_msm_affine_levels(levels))


def set_bucket_slice(slice_index: int = 0, slices: int = 1) -> None:
    """restrict this thread's single-device MSMs to bucket slice `slice_index` of `slices` (include/algebra_b200.h:
    b200_set_msm_bucket_slice); the `slices` partial results over the same inputs add up to the msm.  (0, 1) = whole MSM"""
    _lib.check(_lib.lib().b200_set_msm_bucket_slice(slice_index, slices))


def window_for(curve: G1Curve | int, n: int) -> int:
    cv = CURVES[curve] if isinstance(curve, int) else curve
    return _lib.lib().b200_msm_window_for(cv.cid, n)


def last_timings() -> dict:
    ms = (ctypes.c_float * 7)()
    c, w, adds = ctypes.c_int(), ctypes.c_int(), ctypes.c_ulonglong()
    _lib.lib().b200_msm_last_timings(ms, ctypes.byref(c), ctypes.byref(w), ctypes.byref(adds))
    names = ["digits_hist", "scan", "scatter", "accumulate", "reduce", "combine", "total"]
    d = {k: float(v) for k, v in zip(names, ms)}
    d.update(c=c.value, windows=w.value, bucket_adds=adds.value)
    return d


class HashMapPippenger:
    """stream_pippenger.rs:68-128: coalesce repeated bases (scalars of equal bases are added in Fr on the host) and
    flush through msm_bigint every `buf_size` distinct bases."""

    def __init__(self, curve: G1Curve | int, max_msm_buffer: int):
        self.cv = CURVES[curve] if isinstance(curve, int) else curve
        self.buf_size = max_msm_buffer
        self.buffer: dict[bytes, int] = {}
        self._stream = None

    def add(self, base, scalar):
        """`scalar` is an Fr element as 4 Montgomery limbs (like G::ScalarField)"""
        key = np.asarray(base, dtype=np.uint64).reshape(2 * self.cv.N).tobytes()
        fr = self.cv.fr
        self.buffer[key] = (self.buffer.get(key, 0) + fr.from_limbs(scalar)) % fr.modulus
        if len(self.buffer) == self.buf_size:
            self._flush()

    def _flush(self):
        bases = np.frombuffer(b"".join(self.buffer.keys()), dtype=np.uint64).reshape(-1, 2 * self.cv.N)
        bigints = np.array([[(v >> (64 * i)) & ((1 << 64) - 1) for i in range(4)] for v in self.buffer.values()], dtype=np.uint64)
        if self._stream is None:
            self._stream = MsmStream(self.cv, self.buf_size, 0, _lib.SCALARS_BIGINT)
        self._stream.push(bases, bigints)      # `s.into_bigint()` = canonical limbs
        self.buffer = {}

    def finalize(self) -> np.ndarray:
        if self.buffer:
            self._flush()
        if self._stream is None:
            return msm_unchecked(self.cv, np.zeros((0, 2 * self.cv.N), np.uint64), np.zeros((0, 4), np.uint64))
        st, self._stream = self._stream, None
        return st.finish()
