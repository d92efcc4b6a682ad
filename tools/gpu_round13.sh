#!/bin/bash
for args in "--affine-levels -1" "--affine-levels 4" "--affine-levels 3" "--affine-levels 0" "--window 18 --affine-levels -1" "--window 19 --affine-levels -1" "--window 20 --affine-levels -1"; do timeout 100 python bench.py --steps 3 --warmup 2 --log-n-msm 23 --log-n-ntt 16 --no-cpu-baseline --no-e2e --no-verify $args 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$args', 'c', d['config']['window_c'], round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['phases_ms'].items()})"; done
