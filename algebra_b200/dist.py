"""Multi-GPU MSM: input-chunk or bucket-slice sharding + one exchange of a partial point per rank.

sum_i s_i*P_i splits over any partition of i, so each rank runs the complete single-GPU MSM on its contiguous chunk
and the only communication is an all-gather of one Jacobian point (3N u64 = 144 B for BLS12-381) per rank over
NCCL/NVLink, followed by k-1 point additions on every rank — NCCL cannot reduce curve points, so all-gather + local
sum *is* the all-reduce.  The NTT is not sharded ("replicas only").

Two partitions of the same sum:
  * input chunks (`msm_sharded`, `msm_global`): rank r owns the pairs [lo_r, hi_r) — nothing is replicated, the right split
    when the inputs start on the host (each GPU receives 1/world of the bytes); every rank repeats the per-bucket reduction;
  * bucket slices (`msm_bucket_sliced`): every rank holds ALL pairs (an SRS replicated in every GPU's HBM, scalars produced on
    or broadcast to the devices) and owns the buckets [nb*r/world, nb*(r+1)/world) of every window — sort, accumulation AND
    bucket reduction all shrink by 1/world, only the digit extraction is repeated.

One process per GPU, `torch.distributed` for the plumbing (backend nccl on GPUs; gloo works for the host logic).
`local_msm` / `sum_fn` default to the CUDA library; tests inject CPU stand-ins to exercise the plumbing without a GPU."""
from __future__ import annotations

import numpy as np

from .params import CURVES, G1Curve


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous chunk [lo, hi) of rank `rank`; chunks differ by at most one element and cover [0, n)"""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_points(xyz: np.ndarray, group=None, device=None) -> np.ndarray:
    """(3N,) uint64 partial point of this rank -> (world, 3N) uint64 on every rank"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.from_numpy(np.ascontiguousarray(xyz, dtype=np.uint64).view(np.int64))
    if device is None and dist.get_backend(group) == "nccl":
        device = torch.device("cuda", torch.cuda.current_device())
    if device is not None:
        mine = mine.to(device)
    buf = torch.empty((world, mine.numel()), dtype=torch.int64, device=mine.device)
    dist.all_gather_into_tensor(buf.view(-1), mine, group=group)
    return buf.cpu().numpy().view(np.uint64)


def msm_sharded(curve: G1Curve | int, local_bases, local_scalars, group=None, local_msm=None, sum_fn=None) -> np.ndarray:
    """MSM over the union of all ranks' (bases, scalars) shards; every rank returns the same Projective limbs."""
    import torch.distributed as dist
    from . import variable_base as VB
    cv = CURVES[curve] if isinstance(curve, int) else curve
    local_msm = local_msm or (lambda b, s: VB.msm_unchecked(cv, b, s))
    sum_fn = sum_fn or (lambda pts: VB.sum_points(cv, pts))
    partial = local_msm(local_bases, local_scalars)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return partial
    device = None
    if type(local_bases).__module__.startswith("torch") and local_bases.is_cuda:
        device = local_bases.device
    return sum_fn(all_gather_points(partial, group, device))


def msm_bucket_sliced(curve: G1Curve | int, bases, scalars, group=None, local_msm=None, sum_fn=None) -> np.ndarray:
    """MSM over (bases, scalars) present in full on every rank; rank r computes bucket slice r of world (all windows), the
    partial points are all-gathered and summed.  `local_msm(bases, scalars, slice, slices)` defaults to the CUDA library."""
    import torch.distributed as dist
    from . import variable_base as VB
    cv = CURVES[curve] if isinstance(curve, int) else curve
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1

    def cuda_slice(b, s, i, k):
        VB.set_bucket_slice(i, k)
        try:
            return VB.msm_unchecked(cv, b, s)
        finally:
            VB.set_bucket_slice(0, 1)

    partial = (local_msm or cuda_slice)(bases, scalars, rank, world)
    if world == 1:
        return partial
    sum_fn = sum_fn or (lambda pts: VB.sum_points(cv, pts))
    device = None
    if type(bases).__module__.startswith("torch") and bases.is_cuda:
        device = bases.device
    return sum_fn(all_gather_points(partial, group, device))


def msm_global(curve: G1Curve | int, bases, scalars, group=None, **kw) -> np.ndarray:
    """Same, starting from the full host arrays present on every rank: each rank takes its shard_range chunk."""
    import torch.distributed as dist
    cv = CURVES[curve] if isinstance(curve, int) else curve
    n = min(len(bases), len(scalars))
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_range(n, rank, world)
    return msm_sharded(cv, bases[lo:hi], scalars[lo:hi], group, **kw)
