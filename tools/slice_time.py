"""Time one bucket slice of an MSM on one GPU: what one rank of an N-way bucket-sliced MSM does (tools; not a test).
usage: python tools/slice_time.py LOG_N SLICES [SLICE ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import algebra_b200 as ab
from algebra_b200 import _lib
from algebra_b200 import variable_base as VB

log_n, slices = int(sys.argv[1]), int(sys.argv[2])
which = [int(x) for x in sys.argv[3:]] or [0, slices - 1]
n = 1 << log_n
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
d_bases = torch.empty((n, 12), dtype=torch.int64, device="cuda")
d_b = torch.empty((n,), dtype=torch.int64, device="cuda")
d_s = torch.empty((n, 4), dtype=torch.int64, device="cuda")
_lib.check(L.b200_gen_bases_dev(0, 1, n, d_bases.data_ptr(), d_b.data_ptr(), st))
_lib.check(L.b200_gen_scalars_dev(0, 2, n, d_s.data_ptr(), st))
for i in which:
    VB.set_bucket_slice(i, slices)
    for _ in range(int(os.environ.get('SLICE_TIME_WARM', '2'))):
        ab.msm(0, d_bases, d_s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    reps = int(os.environ.get('SLICE_TIME_REPS', '3'))
    for _ in range(reps):
        ab.msm(0, d_bases, d_s)
    e1.record()
    torch.cuda.synchronize()
    t = VB.last_timings()
    print("n=2^%d slice %d/%d: %.2f ms  phases %s" % (log_n, i, slices, e0.elapsed_time(e1) / reps,
                                                    {k: round(v, 1) for k, v in t.items() if k not in ("bucket_adds",)}), flush=True)
VB.set_bucket_slice(0, 1)
