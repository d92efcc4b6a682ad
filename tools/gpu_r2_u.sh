#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export SLICE_TIME_WARM=1 SLICE_TIME_REPS=1
for S in 8 1; do
  timeout 900 ncu --set full --clock-control none -k regex:msm_pair_add2_kernel -s 4 -c 2 -o gpurun_out/u_pair_S$S -f python tools/slice_time.py 26 $S 0 > gpurun_out/u_ncu_S$S.log 2>&1
  ncu -i gpurun_out/u_pair_S$S.ncu-rep --page details > gpurun_out/u_pair_S$S.details.txt 2>&1
done
ls -la gpurun_out/
