#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r01_pytest_gpu_5.log; cat gpurun_out/r01_pytest_gpu_5.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r01_bench4.json 2> gpurun_out/r01_bench4.err
python -c "
import json; d=json.loads(open('gpurun_out/r01_bench4.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','phases_ms','e2e','gpu_launches','clocks','cpu_baseline')}); print(d['config']); print(d['imad_roofline']['frac'], d['ntt']['value'], d['ntt']['e2e'], d['ntt']['imad_roofline']['frac'])"
timeout 200 python bench.py --steps 3 --warmup 2 --curve 1 --log-n-msm 24 --no-cpu-baseline > gpurun_out/r01_bench_bn254_2e24.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/r01_bench_bn254_2e24.json').read().strip().splitlines()[-1]); print('BN254 2^24', {k:d[k] for k in ('value','ms_per_step','phases_ms','e2e')}, d['config']['window_c'], d['config']['verified_vs_sum_identity'])"
for lg in 20 24; do timeout 200 python bench.py --steps 3 --warmup 2 --log-n-msm $lg --log-n-ntt 20 --no-cpu-baseline > gpurun_out/r01_bench_bls_2e$lg.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r01_bench_bls_2e$lg.json').read().strip().splitlines()[-1]); print('BLS 2^$lg', {k:d[k] for k in ('value','ms_per_step','phases_ms','e2e')}, d['config']['window_c'], d['config']['verified_vs_sum_identity'], 'ntt2^20', d['ntt']['value'])"; done
