#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/t_waves.log
for TW in 1 2 3 4 6 8; do echo "== target_waves=$TW" >> $O; B200_MSM_TARGET_WAVES=$TW timeout 600 python tools/slice_time.py 26 8 0 >> $O 2>&1; done
for TW in 4 8; do echo "== target_waves=$TW wave_min_batch=96" >> $O; B200_MSM_WAVE_MIN_BATCH=96 B200_MSM_TARGET_WAVES=$TW timeout 600 python tools/slice_time.py 26 8 0 >> $O 2>&1; done
for LN in 20 22 23 24; do for TW in 1 3 6; do echo "== n=2^$LN target_waves=$TW" >> $O; B200_MSM_TARGET_WAVES=$TW timeout 600 python tools/slice_time.py $LN 1 0 >> $O 2>&1; done; done
for TW in 3 6; do echo "== S=2 target_waves=$TW" >> $O; B200_MSM_TARGET_WAVES=$TW timeout 600 python tools/slice_time.py 26 2 0 >> $O 2>&1; done
cat $O
