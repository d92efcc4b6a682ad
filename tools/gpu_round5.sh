#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r01_pytest_gpu_6.log; cat gpurun_out/r01_pytest_gpu_6.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench5.json 2> gpurun_out/r01_bench5.err
python -c "
import json; d=json.loads(open('gpurun_out/r01_bench5.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','phases_ms','e2e','gpu_launches')}); print(d['ntt']['value'], d['ntt']['e2e'])"
