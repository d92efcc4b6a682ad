#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-ntt"
for lv in -1 4 3; do echo "== BN254 2^24 levels=$lv" >> gpurun_out/k_levels.log; timeout 300 $B --curve 1 --log-n-msm 24 --affine-levels $lv >> gpurun_out/k_levels.log 2>&1; done
for lv in -1 3 4; do echo "== G2 2^22 levels=$lv" >> gpurun_out/k_levels.log; timeout 600 $B --curve 2 --log-n-msm 22 --affine-levels $lv >> gpurun_out/k_levels.log 2>&1; done
grep -E "^==|ms_per_step" gpurun_out/k_levels.log | cut -c1-330
for plan in "2,7,16" "1,5,16" "2,6,11,16" "3,9,16" "1,3,8,16"; do
  echo "== chunks $plan" >> gpurun_out/k_chunks.log
  B200_MSM_CHUNKS=$plan timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-ntt --no-verify >> gpurun_out/k_chunks.log 2>&1
done
grep -E "^==|ms_per_step" gpurun_out/k_chunks.log | python3 -c "
import sys,json
lab=None
for l in sys.stdin:
    if l.startswith('=='): lab=l.strip(); continue
    try:
        d=json.loads(l); print(lab, 'device', round(d['ms_per_step'],1), 'e2e pinned', round(d['e2e']['ms_per_step'],1), 'pageable', round(d['e2e']['pageable']['ms_per_step'],1))
    except Exception as e: print(lab,'ERR',l[:200])
"
