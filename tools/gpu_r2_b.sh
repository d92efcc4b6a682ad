#!/bin/bash
# round-2 GPU batch B: IMAD peak, NTT gen-2 + G2 parity, NTT timing, pair-add launch lists and ncu captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
./tools/imad_peak > gpurun_out/b_imad_peak.jsonl 2>&1
timeout 1200 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_g2.py -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -15 gpurun_out/b_pytest.log
for g in 1 2; do for ln in 20 22 24 26; do B200_NTT_GENERATION=$g timeout 300 python tools/ntt_time.py --log-n $ln >> gpurun_out/b_ntt_time.jsonl 2>&1; done; done
cat gpurun_out/b_ntt_time.jsonl
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-ntt"
run() { name=$1; shift; echo "== $name" >> gpurun_out/b_variants.log; env B200_POOL_RETAIN_GB=150 "$@" timeout 600 $B >> gpurun_out/b_variants.log 2>&1; }
run v1_nogroup_perthread B200_MSM_PAIR_VARIANT=1 B200_MSM_LEVEL_BUDGET_GB=200 B200_MSM_SHARED_INV=0
run v1_nogroup_shared B200_MSM_PAIR_VARIANT=1 B200_MSM_LEVEL_BUDGET_GB=200
run v2_nogroup_shared B200_MSM_PAIR_VARIANT=2 B200_MSM_LEVEL_BUDGET_GB=200
run v1_budget48_shared B200_MSM_PAIR_VARIANT=1 B200_MSM_LEVEL_BUDGET_GB=48
run v1_budget24_shared B200_MSM_PAIR_VARIANT=1 B200_MSM_LEVEL_BUDGET_GB=24
run v1_budget24_shared_min64 B200_MSM_PAIR_VARIANT=1 B200_MSM_LEVEL_BUDGET_GB=24 B200_MSM_MIN_BATCH=64
run v2_budget24_shared B200_MSM_PAIR_VARIANT=2 B200_MSM_LEVEL_BUDGET_GB=24
grep -E "^==|ms_per_step" gpurun_out/b_variants.log | cut -c1-330
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
BL="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-ntt --no-verify"
B200_POOL_RETAIN_GB=150 B200_MSM_LEVEL_BUDGET_GB=200 B200_MSM_PAIR_VARIANT=1 timeout 600 $NCU --log-file gpurun_out/b_launch_v1.csv $BL > gpurun_out/b_launch_v1.log 2>&1
B200_POOL_RETAIN_GB=150 B200_MSM_LEVEL_BUDGET_GB=200 B200_MSM_PAIR_VARIANT=2 timeout 600 $NCU --log-file gpurun_out/b_launch_v2.csv $BL > gpurun_out/b_launch_v2.log 2>&1
# full captures of the two dominant pair-add launches (level 1, level 2) at a 1/4-scale replica of the metric config
# (n = 2^24, c = 18: 128 entries per bucket like n = 2^26, c = 20), and of the three NTT passes at n = 2^24
FULL="ncu --set full --import-source on --clock-control none"
BS="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-ntt --no-verify --log-n-msm 24 --window 18"
B200_MSM_LEVEL_BUDGET_GB=200 B200_MSM_PAIR_VARIANT=1 timeout 900 $FULL -k regex:msm_pair_add -c 2 -o gpurun_out/b_pair_v1 -f $BS > gpurun_out/b_ncu_v1.log 2>&1
B200_MSM_LEVEL_BUDGET_GB=200 B200_MSM_PAIR_VARIANT=2 timeout 900 $FULL -k regex:msm_pair_add -c 2 -o gpurun_out/b_pair_v2 -f $BS > gpurun_out/b_ncu_v2.log 2>&1
timeout 600 $FULL -k regex:ntt2_pass -c 3 -o gpurun_out/b_ntt2 -f python tools/ntt_time.py --log-n 24 --reps 1 > gpurun_out/b_ncu_ntt2.log 2>&1
ls -la gpurun_out/
