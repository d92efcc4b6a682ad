#!/usr/bin/env python3
"""b200_msm_sw_g1_multi / b200_bases_upload + b200_msm_bases on 1, 2, 4, 8 devices from ONE process (no torch.distributed):
n = 2^26 pairs in host memory (pinned and pageable), wall-clock per call, result checked against the 1-device result."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import algebra_b200 as ab
from algebra_b200 import _lib, variable_base as VB

L = _lib.lib()
log_n = int(os.environ.get("LOG_N", "26"))
n = 1 << log_n
torch.cuda.set_device(0)
st = torch.cuda.current_stream().cuda_stream
d_bases = torch.empty((n, 12), dtype=torch.int64, device="cuda")
d_s = torch.empty((n, 4), dtype=torch.int64, device="cuda")
_lib.check(L.b200_gen_bases_dev(0, 1, n, d_bases.data_ptr(), None, st))
_lib.check(L.b200_gen_scalars_dev(0, 2, n, d_s.data_ptr(), st))
hb = torch.empty((n, 12), dtype=torch.int64).pin_memory(); hb.copy_(d_bases)
hs = torch.empty((n, 4), dtype=torch.int64).pin_memory(); hs.copy_(d_s)
del d_bases, d_s
torch.cuda.empty_cache()
pin_b, pin_s = hb.numpy().view(np.uint64), hs.numpy().view(np.uint64)
ref = None
for ng in [g for g in (1, 2, 4, 8) if g <= VB.device_count()]:
    for kind, (b, s) in (("pinned", (pin_b, pin_s)),):
        VB.msm_multi(0, b, s, ng)
        t0 = time.perf_counter(); r = VB.msm_multi(0, b, s, ng); dt = time.perf_counter() - t0
        aff = ab.into_affine(0, r)
        ref = aff if ref is None else ref
        print(json.dumps({"entry": "b200_msm_sw_g1_multi", "ngpus": ng, "host_memory": kind, "log_n": log_n, "ms": dt * 1e3, "MSM_per_s": 1 / dt,
                          "same_as_1gpu": bool((aff == ref).all())}), flush=True)
    h = VB.bases_upload(0, pin_b, ng)
    VB.msm_with_bases(h, pin_s)
    t0 = time.perf_counter(); r = VB.msm_with_bases(h, pin_s); dt = time.perf_counter() - t0
    print(json.dumps({"entry": "b200_msm_bases (resident bases)", "ngpus": ng, "log_n": log_n, "ms": dt * 1e3, "MSM_per_s": 1 / dt,
                      "same_as_1gpu": bool((ab.into_affine(0, r) == ref).all())}), flush=True)
    VB.bases_free(h)
pb, ps = pin_b.copy(), pin_s.copy()     # pageable
ng = VB.device_count()
VB.msm_multi(0, pb, ps, ng)
t0 = time.perf_counter(); r = VB.msm_multi(0, pb, ps, ng); dt = time.perf_counter() - t0
print(json.dumps({"entry": "b200_msm_sw_g1_multi", "ngpus": ng, "host_memory": "pageable", "log_n": log_n, "ms": dt * 1e3, "MSM_per_s": 1 / dt,
                  "same_as_1gpu": bool((ab.into_affine(0, r) == ref).all())}), flush=True)
