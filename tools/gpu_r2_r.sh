#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for S in 1 2 4 8; do timeout 600 python tools/slice_time.py 26 $S >> gpurun_out/r_slices.log 2>&1; done
cat gpurun_out/r_slices.log
