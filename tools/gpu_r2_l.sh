#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/l_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest_gpu.log; tail -4 gpurun_out/l_pytest_gpu.log | cut -c1-200
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-ntt"
for cfg in "" "--log-n-msm 23" "--log-n-msm 24" "--log-n-msm 20" "--log-n-msm 16" "--curve 1 --log-n-msm 24" "--curve 2 --log-n-msm 22" "--curve 2 --log-n-msm 20"; do
  echo "== default $cfg" >> gpurun_out/l_bench.log; timeout 600 $B $cfg >> gpurun_out/l_bench.log 2>&1
done
grep -E "^==|ms_per_step" gpurun_out/l_bench.log | python3 -c "
import sys,json
lab=None
for l in sys.stdin:
    if l.startswith('=='): lab=l.strip(); continue
    try:
        d=json.loads(l); print(lab, round(d['ms_per_step'],2), {k:round(v,1) for k,v in d['phases_ms'].items()}, 'c', d['window_c'], d.get('verified'))
    except Exception as e: print(lab,'ERR',l[:200])
"
