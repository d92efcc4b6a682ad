"""N > 1 host logic on CPU: world_size-2 (and 3) gloo process groups run algebra_b200.dist's shard -> partial ->
all-gather -> sum path with CPU stand-ins for the two GPU calls (the oracle's restated reference MSM and Jacobian
addition), and every rank must end with the same point as the unsharded MSM."""
import os
import random
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import coracle as C
from oracle import pyoracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, bases, scalars, want, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from algebra_b200 import dist as D

    def cpu_msm(b, s):
        if len(b) == 0:
            z = np.zeros(18, dtype=np.uint64)
            z[:6] = O.BLS12_381_FQ.limbs(O.BLS12_381_FQ.R)
            z[6:12] = O.BLS12_381_FQ.limbs(O.BLS12_381_FQ.R)
            return z
        return C.msm(0, b, s, threads=1)

    def cpu_sum(pts):
        acc = pts[0:1].copy()
        for k in range(1, pts.shape[0]):
            acc = C.ec_op(0, "jac_add", acc, pts[k:k + 1])
        return acc.reshape(-1)

    xyz = D.msm_global(0, bases, scalars, local_msm=cpu_msm, sum_fn=cpu_sum)
    aff = C.ec_op(0, "jac_to_affine", xyz.reshape(1, -1)).reshape(-1)
    q.put((rank, bool((aff == want).all())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 257), (3, 64), (2, 1)])
def test_sharded_msm_gloo(world, n):
    cv = O.BLS12_381
    rnd = random.Random(n)
    ks = [rnd.randrange(1, 1 << 30) for _ in range(n)]
    bases = cv.encode_affine([cv.mul(cv.G, k) for k in ks])
    sc = [rnd.randrange(cv.fr.p) for _ in range(n)]
    scalars = cv.fr.encode(sc)
    want = cv.encode_affine([cv.mul(cv.G, sum(k * s for k, s in zip(ks, sc)) % cv.fr.p)])[0]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, bases, scalars, want, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


def _slice_worker(rank, world, port, ks, sc, bases, scalars, want, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from algebra_b200 import dist as D
    cv = O.BLS12_381
    c = 5
    W = (255 + c - 1) // c

    def cpu_slice(b, s, i, k):
        """signed-digit Pippenger restricted to the magnitudes whose bucket lies in slice i of k (the same cut as make_geom:
        [nb*i/k, nb*(i+1)/k) per window, top window unsigned); bases are known multiples of G, so the slice's sum is one scalar"""
        assert len(b) == len(bases) and len(s) == len(scalars)      # every rank is handed the full inputs
        tot = 0
        for kb, v in zip(ks, sc):
            carry = 0
            for w in range(W):
                d = ((v >> (w * c)) & ((1 << c) - 1)) + carry
                top = w == W - 1
                nbw = 1 << (255 - (W - 1) * c) if top else 1 << (c - 1)
                carry = 0
                if not top and d >= (1 << (c - 1)):
                    d -= 1 << c
                    carry = 1
                if d and nbw * i // k <= abs(d) - 1 < nbw * (i + 1) // k:
                    tot += d * kb << (w * c)
        return _jac(cv, cv.mul(cv.G, tot % cv.fr.p))

    def cpu_sum(pts):
        acc = pts[0:1].copy()
        for j in range(1, pts.shape[0]):
            acc = C.ec_op(0, "jac_add", acc, pts[j:j + 1])
        return acc.reshape(-1)

    xyz = D.msm_bucket_sliced(0, bases, scalars, local_msm=cpu_slice, sum_fn=cpu_sum)
    aff = C.ec_op(0, "jac_to_affine", xyz.reshape(1, -1)).reshape(-1)
    q.put((rank, bool((aff == want).all())))
    dist.barrier()
    dist.destroy_process_group()


def _jac(cv, pt):
    """affine point (or None) -> Projective limbs (x, y, 1) / (1, 1, 0) in Montgomery form"""
    fq = cv.fq
    out = np.zeros(3 * fq.N, dtype=np.uint64)
    one = fq.limbs(fq.R)
    if pt is None:
        out[:fq.N], out[fq.N:2 * fq.N] = one, one
        return out
    out[:2 * fq.N] = cv.encode_affine([pt])[0]
    out[2 * fq.N:] = one
    return out


@pytest.mark.parametrize("world", [2, 3])
def test_bucket_sliced_msm_gloo(world):
    cv = O.BLS12_381
    rnd = random.Random(world)
    n = 40
    ks = [rnd.randrange(1, 1 << 30) for _ in range(n)]
    bases = cv.encode_affine([cv.mul(cv.G, k) for k in ks])
    sc = [rnd.randrange(cv.fr.p) for _ in range(n)]
    scalars = cv.fr.encode(sc)
    want = cv.encode_affine([cv.mul(cv.G, sum(k * s for k, s in zip(ks, sc)) % cv.fr.p)])[0]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slice_worker, args=(r, world, port, ks, sc, bases, scalars, want, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


def test_shard_range_partition():
    from algebra_b200.dist import shard_range
    for n in (0, 1, 7, 64, 1 << 26, (1 << 26) + 5):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
