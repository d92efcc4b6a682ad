#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msm.py -x -q -k "slice or window_sweep or chunked or large_sizes" > gpurun_out/q_pytest_slices.log 2>&1; tail -3 gpurun_out/q_pytest_slices.log
for S in 2 4 8; do timeout 600 python tools/slice_time.py 26 $S >> gpurun_out/q_slices.log 2>&1; done
cat gpurun_out/q_slices.log
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-ntt"
echo "== pad=0" >> gpurun_out/p_variants.log; timeout 600 $B >> gpurun_out/p_variants.log 2>&1
echo "== pad=1" >> gpurun_out/p_variants.log; B200_MSM_PAD_BASES=1 timeout 600 $B >> gpurun_out/p_variants.log 2>&1
echo "== pad=1 2^23" >> gpurun_out/p_variants.log; B200_MSM_PAD_BASES=1 timeout 600 $B --log-n-msm 23 >> gpurun_out/p_variants.log 2>&1
for rep in 1 2; do echo "== G2 2^20 rep $rep" >> gpurun_out/p_variants.log; timeout 600 $B --curve 2 --log-n-msm 20 >> gpurun_out/p_variants.log 2>&1; done
grep -E "^==|ms_per_step" gpurun_out/p_variants.log | python3 -c "
import sys,json
lab=None
for l in sys.stdin:
    if l.startswith('=='): lab=l.strip(); continue
    try:
        d=json.loads(l); print(lab, round(d['ms_per_step'],2), {k:round(v,1) for k,v in d['phases_ms'].items()}, d.get('verified'))
    except Exception as e: print(lab,'ERR',l[:200])
"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest_all.log 2>&1; tail -3 gpurun_out/q_pytest_all.log
