"""Multi-rank parity on real GPUs: one process per GPU over NCCL, input-chunk sharding, one all-gather of the
partial points, local Jacobian sum — every rank must hold the same group element as the unsharded identity
MSM(b_i*G, s_i) = (sum s_i*b_i mod r)*G.  Skipped on boxes with fewer than 2 GPUs; also covers the single-process
multi-device C-ABI entry point b200_msm_sw_g1_multi (one host thread and one stream per device, no NCCL)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _expected(cid, b_host, s_host):
    from oracle import pyoracle as O
    cv = O.CURVES[cid]
    sv = cv.fr.decode(s_host)
    tot = sum(int(b) * s for b, s in zip(b_host, sv)) % cv.fr.p
    return cv.encode_affine([cv.mul(cv.G, tot)])[0]


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import algebra_b200 as ab
    from algebra_b200 import _lib
    from algebra_b200 import dist as D
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    d_bases = torch.empty((n, 12), dtype=torch.int64, device="cuda")
    d_b = torch.empty((n,), dtype=torch.int64, device="cuda")
    d_s = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    _lib.check(L.b200_gen_bases_dev(0, 4711, n, d_bases.data_ptr(), d_b.data_ptr(), st))    # same seed on every rank
    _lib.check(L.b200_gen_scalars_dev(0, 4712, n, d_s.data_ptr(), st))
    lo, hi = D.shard_range(n, rank, world)
    xyz = D.msm_sharded(0, d_bases[lo:hi].contiguous(), d_s[lo:hi].contiguous())
    got = ab.into_affine(0, xyz)
    want = _expected(0, d_b.cpu().numpy().view(np.uint64), d_s.cpu().numpy().view(np.uint64))
    q.put((rank, bool((got == want).all())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1 << 14, 4097])
def test_sharded_msm_nccl_world2(n):
    if _ngpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    assert sorted(res) == [(0, True), (1, True)]


@pytest.mark.parametrize("ngpus", [1, 2])
def test_multi_device_c_abi(ngpus):
    """b200_msm_sw_g1_multi: host buffers in, one result out, `ngpus` devices driven from ONE process through the header
    only.  ngpus = 1 runs on any box; ngpus = 2 needs two devices."""
    if _ngpus() < ngpus:
        pytest.skip("needs >= %d GPUs" % ngpus)
    import torch
    import algebra_b200 as ab
    from algebra_b200 import _lib
    from algebra_b200 import variable_base as VB
    L = _lib.lib()
    n = (1 << 15) + 13
    torch.cuda.set_device(0)
    st = torch.cuda.current_stream().cuda_stream
    d_bases = torch.empty((n, 12), dtype=torch.int64, device="cuda")
    d_b = torch.empty((n,), dtype=torch.int64, device="cuda")
    d_s = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    _lib.check(L.b200_gen_bases_dev(0, 99, n, d_bases.data_ptr(), d_b.data_ptr(), st))
    _lib.check(L.b200_gen_scalars_dev(0, 98, n, d_s.data_ptr(), st))
    bh, sh = d_bases.cpu().numpy().view(np.uint64), d_s.cpu().numpy().view(np.uint64)
    want = _expected(0, d_b.cpu().numpy().view(np.uint64), sh)
    got = ab.into_affine(0, VB.msm_multi(0, bh, sh, ngpus))
    assert (got == want).all()
    VB.set_bucket_slice(1, 2)      # input-chunk sharding computes whole MSMs: a bucket slice set on the thread is ignored ...
    try:
        assert (ab.into_affine(0, VB.msm_multi(0, bh, sh, ngpus)) == want).all()
        assert not (ab.into_affine(0, ab.msm(0, bh, sh)) == want).all()   # ... and still in force afterwards for the single-device call
    finally:
        VB.set_bucket_slice(0, 1)
    # resident bases: upload once (sharded over the devices), reuse for two scalar vectors
    h = VB.bases_upload(0, bh, ngpus)
    try:
        assert (ab.into_affine(0, VB.msm_with_bases(h, sh)) == want).all()
        sh2 = sh[::-1].copy()
        want2 = _expected(0, d_b.cpu().numpy().view(np.uint64), sh2)
        assert (ab.into_affine(0, VB.msm_with_bases(h, sh2)) == want2).all()
        # fewer scalars than bases: Err(min_len) in the mirror; the C ABI takes n <= handle size
        assert (ab.into_affine(0, VB.msm_with_bases(h, sh[:1000])) == _expected(0, d_b.cpu().numpy().view(np.uint64)[:1000], sh[:1000])).all()
    finally:
        VB.bases_free(h)


def _worker_slices(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import algebra_b200 as ab
    from algebra_b200 import _lib
    from algebra_b200 import dist as D
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    d_bases = torch.empty((n, 12), dtype=torch.int64, device="cuda")
    d_b = torch.empty((n,), dtype=torch.int64, device="cuda")
    d_s = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    _lib.check(L.b200_gen_bases_dev(0, 4811, n, d_bases.data_ptr(), d_b.data_ptr(), st))    # replicated inputs: same seed on every rank
    _lib.check(L.b200_gen_scalars_dev(0, 4812, n, d_s.data_ptr(), st))
    xyz = D.msm_bucket_sliced(0, d_bases, d_s)
    got = ab.into_affine(0, xyz)
    want = _expected(0, d_b.cpu().numpy().view(np.uint64), d_s.cpu().numpy().view(np.uint64))
    q.put((rank, bool((got == want).all())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1 << 14, 4097])
def test_bucket_sliced_msm_nccl_world2(n):
    """bucket-slice sharding over NCCL: every rank holds all pairs and reduces its own range of every window's buckets"""
    if _ngpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_slices, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
