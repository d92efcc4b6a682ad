#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r01_pytest_gpu_7.log; cat gpurun_out/r01_pytest_gpu_7.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench6.json 2> gpurun_out/r01_bench6.err
python -c "
import json; d=json.loads(open('gpurun_out/r01_bench6.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','phases_ms','e2e','gpu_launches')}); print(d['ntt']['value'], d['ntt']['e2e'])"
for g in 0 1 3; do B200_NTT_LOGG=$g timeout 120 python bench.py --steps 5 --warmup 3 --log-n-msm 16 --no-cpu-baseline --no-e2e --no-verify 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('logG=$g', d['ntt']['ms_per_step'], d['ntt']['roundtrip_ok'])"; done
