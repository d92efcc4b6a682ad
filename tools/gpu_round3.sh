#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r01_pytest_gpu_4.log; cat gpurun_out/r01_pytest_gpu_4.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench3.json 2> gpurun_out/r01_bench3.err
python -c "
import json; d=json.loads(open('gpurun_out/r01_bench3.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','phases_ms','e2e','gpu_launches','imad_roofline')}); print(d['ntt']['value'], d['ntt']['e2e'])"
rm -f gpurun_out/r01_window_sweep2.jsonl
for c in 16 18 19 21 22 23; do timeout 120 python bench.py --steps 2 --warmup 1 --window $c --no-e2e --no-cpu-baseline --no-verify 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'c':d['config']['window_c'],'W':d['config']['windows'],'ms':d['ms_per_step'],'phases':d['phases_ms']}))" >> gpurun_out/r01_window_sweep2.jsonl; done
cat gpurun_out/r01_window_sweep2.jsonl
