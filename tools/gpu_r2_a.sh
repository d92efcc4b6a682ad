#!/bin/bash
# round-2 GPU batch A: full -m gpu suite on the refactored MSM, then pair-add / level-budget / L2-fetch variants at 2^26
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-ntt"
run() { name=$1; shift; echo "== $name" >> gpurun_out/a_variants.log; env "$@" timeout 600 $B >> gpurun_out/a_variants.log 2>&1; }
run v2_default X=1
run v1_default B200_MSM_PAIR_VARIANT=1
run v2_nogroup B200_MSM_LEVEL_BUDGET_GB=200
run v1_nogroup B200_MSM_PAIR_VARIANT=1 B200_MSM_LEVEL_BUDGET_GB=200
run v2_l2gran32 B200_L2_FETCH_GRANULARITY=32
run v2_l2gran128 B200_L2_FETCH_GRANULARITY=128
run v2_budget8 B200_MSM_LEVEL_BUDGET_GB=8
grep -E "^==|ms_per_step" gpurun_out/a_variants.log | cut -c1-400
