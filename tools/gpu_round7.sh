#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r01_pytest_gpu_8.log; cat gpurun_out/r01_pytest_gpu_8.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench7.json 2> gpurun_out/r01_bench7.err
python -c "
import json; d=json.loads(open('gpurun_out/r01_bench7.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','phases_ms','e2e','gpu_launches')}); print(d['ntt']['value'], d['ntt']['ms_per_step'], d['ntt']['ifft_ms_per_step'], d['ntt']['e2e'])"
