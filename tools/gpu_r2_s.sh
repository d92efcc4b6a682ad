#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export SLICE_TIME_WARM=1 SLICE_TIME_REPS=1
for S in 1 8; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s_launches_S$S.csv python tools/slice_time.py 26 $S 0 > gpurun_out/s_ncu_S$S.log 2>&1
  python tools/launch_sum.py gpurun_out/s_launches_S$S.csv
done
