#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/g2_debug.py > gpurun_out/d_g2_debug.log 2>&1; cat gpurun_out/d_g2_debug.log
cat > /tmp/g2_tiny.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, algebra_b200 as ab
from oracle import pyoracle as O
G2, fr = O.BLS12_381_G2, O.BLS12_381_FR
P = [G2.mul(G2.G, k) for k in (5, 9)]
print(ab.into_affine(2, ab.msm(2, G2.encode_affine(P), fr.encode([3, 2])))[:2])
PY
timeout 900 compute-sanitizer --tool memcheck python /tmp/g2_tiny.py > gpurun_out/d_g2_memcheck.log 2>&1; tail -15 gpurun_out/d_g2_memcheck.log
timeout 900 compute-sanitizer --tool initcheck python /tmp/g2_tiny.py > gpurun_out/d_g2_initcheck.log 2>&1; tail -15 gpurun_out/d_g2_initcheck.log
for g in 1 2; do for ln in 20 22 24 26; do B200_NTT_GENERATION=$g timeout 300 python tools/ntt_time.py --log-n $ln >> gpurun_out/d_ntt_time.jsonl 2>&1; done; done
cat gpurun_out/d_ntt_time.jsonl
timeout 900 python -m pytest tests/test_gpu_ntt.py -x -q > gpurun_out/d_pytest_ntt.log 2>&1; tail -3 gpurun_out/d_pytest_ntt.log
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-ntt"
echo "== default (all gen-2, cp.async.ca)" >> gpurun_out/d_variants.log; timeout 600 $B >> gpurun_out/d_variants.log 2>&1
echo "== per-thread inversion" >> gpurun_out/d_variants.log; B200_MSM_SHARED_INV=0 timeout 600 $B >> gpurun_out/d_variants.log 2>&1
grep -E "^==|ms_per_step" gpurun_out/d_variants.log | cut -c1-330
for ln in 22 23 24 25; do for c in 15 16 17 18 20; do
  echo "== n=2^$ln c=$c" >> gpurun_out/d_sweep.log
  timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-ntt --no-verify --log-n-msm $ln --window $c >> gpurun_out/d_sweep.log 2>&1
done; done
grep -E "^==|ms_per_step" gpurun_out/d_sweep.log | cut -c1-200
