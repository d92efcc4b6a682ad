"""Field-operation micro-benchmark on the GPU, the op list of the reference's criterion templates
(bench-templates/src/macros/field.rs:69-155: Addition, Subtraction, Negation, Double, Multiplication, Square, Inverse,
Into/From BigInt) through b200_fp_op_dev (1 thread per element, `reps` dependent applications), with the CPU oracle's
single-thread rate beside it.  Prints JSON lines."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from algebra_b200 import _lib
from oracle import coracle as C

L = _lib.lib()
OPS = [("Multiplication", 0, 64), ("Addition", 1, 256), ("Subtraction", 2, 256), ("Square", 3, 64), ("Double", 4, 256),
       ("Negation", 5, 256), ("Into BigInt", 6, 64), ("From BigInt", 7, 64), ("Inverse", 8, 1)]
NAMES = {0: "bls12_381_fq", 1: "bls12_381_fr", 2: "bn254_fq", 3: "bn254_fr"}
LIMBS = {0: 6, 1: 4, 2: 4, 3: 4}
WIDE = {0: 300, 1: 136, 2: 136, 3: 136}
st = torch.cuda.current_stream().cuda_stream
for fid in (0, 1, 2, 3):
    n = 1 << 22
    N = LIMBS[fid]
    a = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    _lib.check(L.b200_gen_scalars_dev(0 if fid < 2 else 1, 3 + fid, n, a.data_ptr(), st))   # 256-bit values < r < p: valid in all fields
    x = torch.zeros((n, N), dtype=torch.int64, device="cuda")
    x[:, :4] = a
    y = x.roll(1, 0).contiguous()
    out = torch.empty_like(x)
    for name, op, reps in OPS:
        m = n if op != 8 else 1 << 16
        _lib.check(L.b200_fp_op_dev(fid, op, x.data_ptr(), y.data_ptr(), out.data_ptr(), m, reps, st))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.b200_fp_op_dev(fid, op, x.data_ptr(), y.data_ptr(), out.data_ptr(), m, reps, st))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        gops = m * reps / ms / 1e6
        hx = x[:4096].cpu().numpy().view(np.uint64)
        hy = y[:4096].cpu().numpy().view(np.uint64)
        cname = {0: "mul", 1: "add", 2: "sub", 3: "sqr", 4: "dbl", 5: "neg", 6: "into_bigint", 7: "from_bigint", 8: "inv"}[op]
        hx1 = hx[:256] if op == 8 else hx
        t0 = time.perf_counter()
        for _ in range(3):
            C.fp_op(fid, cname, hx1, hy[: len(hx1)])
        cpu_ns = (time.perf_counter() - t0) / 3 / len(hx1) * 1e9
        rec = {"field": NAMES[fid], "op": name, "gpu_Gop_per_s": round(gops, 2), "cpu_oracle_ns_per_op_1thread": round(cpu_ns, 1)}
        if op in (0, 3):
            rec["wide_mad_T_per_s"] = round(gops * WIDE[fid] / 1e3, 2)
        print(json.dumps(rec), flush=True)
