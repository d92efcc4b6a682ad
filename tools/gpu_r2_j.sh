#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-ntt"
run() { name=$1; shift; echo "== $name" >> gpurun_out/j_variants.log; env "$@" timeout 600 $B >> gpurun_out/j_variants.log 2>&1; }
run default X=1
run stagger B200_MSM_STAGGER=1
echo "== 2^23 default" >> gpurun_out/j_variants.log; timeout 300 $B --log-n-msm 23 >> gpurun_out/j_variants.log 2>&1
echo "== 2^23 stagger" >> gpurun_out/j_variants.log; B200_MSM_STAGGER=1 timeout 300 $B --log-n-msm 23 >> gpurun_out/j_variants.log 2>&1
echo "== 2^24 stagger" >> gpurun_out/j_variants.log; B200_MSM_STAGGER=1 timeout 300 $B --log-n-msm 24 >> gpurun_out/j_variants.log 2>&1
grep -E "^==|ms_per_step" gpurun_out/j_variants.log | cut -c1-420
