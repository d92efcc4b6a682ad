#!/bin/bash
mkdir -p gpurun_out
for mb in 5 6; do B200_PAIR_MINB=$mb timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --log-n-ntt 16 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('minb=$mb', d['ms_per_step'], d['phases_ms']['accumulate'], d['config']['verified_vs_sum_identity'])"; done
for c in 19 21 22; do timeout 200 python bench.py --steps 2 --warmup 1 --window $c --no-cpu-baseline --no-e2e --no-verify --log-n-ntt 16 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c=$c', d['ms_per_step'], d['phases_ms'])"; done
