#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r01_pytest_gpu_10.log; cat gpurun_out/r01_pytest_gpu_10.log
rm -f gpurun_out/r01_affine_levels2.jsonl
for lv in 3 4 -1; do timeout 200 python bench.py --steps 2 --warmup 1 --affine-levels $lv --no-cpu-baseline --log-n-ntt 16 2>gpurun_out/r01_affine_err_$lv.log | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'levels':$lv,'c':d['config']['window_c'],'ms':d['ms_per_step'],'verified':d['config']['verified_vs_sum_identity'],'e2e':d['e2e'],'phases':d['phases_ms']}))" >> gpurun_out/r01_affine_levels2.jsonl; done
cat gpurun_out/r01_affine_levels2.jsonl
ncu --set full --clock-control none --import-source on -k regex:msm_pair_add -c 2 -o gpurun_out/r01_prof_pair -f python bench.py --steps 1 --warmup 1 --log-n-msm 22 --log-n-ntt 16 --affine-levels 2 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/r01_ncu_pair.log 2>&1
