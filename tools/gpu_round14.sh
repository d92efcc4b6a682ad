#!/bin/bash
for ch in "2,5,16" "2,7,16" "3,7,16" "2,6,11,16" "2,5,10,16"; do B200_MSM_CHUNKS=$ch timeout 120 python bench.py --steps 2 --warmup 1 --log-n-ntt 16 --no-cpu-baseline --no-verify 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunks $ch', 'e2e ms', round(d['e2e']['ms_per_step'],1), 'dev ms', round(d['ms_per_step'],1), d['e2e']['same_result'])"; done
