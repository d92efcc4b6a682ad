#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-ntt"
echo "== l1_minb=4" >> gpurun_out/o_variants.log; timeout 600 $B >> gpurun_out/o_variants.log 2>&1
echo "== l1_minb=5" >> gpurun_out/o_variants.log; B200_MSM_L1_MINB=5 timeout 600 $B >> gpurun_out/o_variants.log 2>&1
echo "== l1_minb=5 2^23" >> gpurun_out/o_variants.log; B200_MSM_L1_MINB=5 timeout 600 $B --log-n-msm 23 >> gpurun_out/o_variants.log 2>&1
grep -E "^==|ms_per_step" gpurun_out/o_variants.log | python3 -c "
import sys,json
lab=None
for l in sys.stdin:
    if l.startswith('=='): lab=l.strip(); continue
    try:
        d=json.loads(l); print(lab, round(d['ms_per_step'],2), {k:round(v,1) for k,v in d['phases_ms'].items()}, d.get('verified'))
    except Exception as e: print(lab,'ERR',l[:200])
"
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/o_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/o_pytest_gpu.log; tail -4 gpurun_out/o_pytest_gpu.log | cut -c1-200
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err; tail -c 600 gpurun_out/o_bench.json
for cfg in "--log-n-msm 24" "--log-n-msm 20" "--log-n-msm 16" "--curve 1 --log-n-msm 24" "--curve 2 --log-n-msm 22" "--curve 2 --log-n-msm 20"; do
  echo "== $cfg" >> gpurun_out/o_bench_configs.log
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ntt $cfg >> gpurun_out/o_bench_configs.log 2>&1
done
grep -E "^==|ms_per_step" gpurun_out/o_bench_configs.log | python3 -c "
import sys,json
lab=None
for l in sys.stdin:
    if l.startswith('=='): lab=l.strip(); continue
    try:
        d=json.loads(l); print(lab, 'dev', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), round(d['e2e']['pageable']['ms_per_step'],2), 'c', d['window_c'], d['verified'])
    except Exception as e: print(lab,'ERR',l[:200])
"
