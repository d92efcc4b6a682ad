#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_msm.py -m gpu -q -k "affine or kat or random" 2>&1 | tail -15 > gpurun_out/r01_pytest_affine.log; cat gpurun_out/r01_pytest_affine.log
rm -f gpurun_out/r01_affine_levels.jsonl
for lv in 0 1 2 3 4; do timeout 200 python bench.py --steps 2 --warmup 1 --affine-levels $lv --no-e2e --no-cpu-baseline --log-n-ntt 16 2>gpurun_out/r01_affine_err_$lv.log | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'levels':$lv,'c':d['config']['window_c'],'ms':d['ms_per_step'],'verified':d['config']['verified_vs_sum_identity'],'phases':d['phases_ms']}))" >> gpurun_out/r01_affine_levels.jsonl; done
cat gpurun_out/r01_affine_levels.jsonl; tail -3 gpurun_out/r01_affine_err_1.log
