#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -22 > gpurun_out/r01_pytest_gpu_11.log; cat gpurun_out/r01_pytest_gpu_11.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r01_bench8.json 2> gpurun_out/r01_bench8.err
python -c "
import json; d=json.loads(open('gpurun_out/r01_bench8.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','phases_ms','e2e','gpu_launches')}); print(d['config']['verified_vs_sum_identity'], d['ntt']['value'])"
