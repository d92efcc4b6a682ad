#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/field_bench.py > gpurun_out/r01_field_bench.jsonl 2> gpurun_out/r01_field_bench.err; tail -3 gpurun_out/r01_field_bench.err; head -12 gpurun_out/r01_field_bench.jsonl
timeout 200 python bench.py --steps 3 --warmup 2 --curve 1 --log-n-msm 24 --no-cpu-baseline > gpurun_out/r01_bench_bn254_2e24.json 2>/dev/null
for lg in 16 20 24; do timeout 200 python bench.py --steps 3 --warmup 2 --log-n-msm $lg --log-n-ntt 20 --no-cpu-baseline > gpurun_out/r01_bench_bls_2e$lg.json 2>/dev/null; done
for f in bn254_2e24 bls_2e16 bls_2e20 bls_2e24; do python -c "
import json; d=json.loads(open('gpurun_out/r01_bench_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value'],2), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'c', d['config']['window_c'], d['config']['verified_vs_sum_identity'], 'ntt2^20', round(d['ntt']['value'],1))"; done
