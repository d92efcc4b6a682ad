#!/bin/bash
set -x
mkdir -p gpurun_out
cat /sys/fs/cgroup/cpu.max > gpurun_out/r01_cpuinfo.txt 2>&1; nproc >> gpurun_out/r01_cpuinfo.txt; uptime >> gpurun_out/r01_cpuinfo.txt
python - >> gpurun_out/r01_cpuinfo.txt 2>&1 <<'PY'
import time, numpy as np
from oracle import coracle as C, pyoracle as O
print("oracle threads", C.num_threads())
cv=O.BLS12_381
pts=cv.encode_affine([cv.mul(cv.G,k) for k in range(1,65)])
n=1<<20
bases=np.ascontiguousarray(np.tile(pts,(n//64,1)))
rng=np.random.default_rng(1); sc=rng.integers(0,1<<64,size=(n,4),dtype=np.uint64); sc[:,3]&=np.uint64((1<<62)-1)
for t in (1,2,8,16,32,64,128):
    t0=time.perf_counter(); C.msm(0,bases[:n//8 if t==1 else n],sc[:n//8 if t==1 else n],threads=t); print("threads",t,"n",n//8 if t==1 else n,"sec",round(time.perf_counter()-t0,3))
PY
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r01_pytest_gpu_3.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r01_bench2.json 2> gpurun_out/r01_bench2.err
for c in 17 18 19 21 22; do python bench.py --steps 2 --warmup 1 --window $c --no-e2e --no-cpu-baseline --no-verify 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'c':d['config']['window_c'],'W':d['config']['windows'],'ms':d['ms_per_step'],'phases':d['phases_ms']}))" >> gpurun_out/r01_window_sweep.jsonl; done
cat gpurun_out/r01_window_sweep.jsonl; cat gpurun_out/r01_cpuinfo.txt; cat gpurun_out/r01_pytest_gpu_3.log
python -c "
import json; d=json.loads(open('gpurun_out/r01_bench2.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','phases_ms','e2e','cpu_baseline')}); print(d['ntt'])"
