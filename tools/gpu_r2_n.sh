#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/n_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/n_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/n_sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/n_sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool initcheck python tools/sanitize_small.py > gpurun_out/n_sanitizer_initcheck.log 2>&1; tail -3 gpurun_out/n_sanitizer_initcheck.log
B200_NTT_GENERATION=2 B200_MSM_PAIR_VARIANT=1 B200_MSM_PAIR_VARIANT_L1=1 timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/n_sanitizer_memcheck_alt.log 2>&1; tail -3 gpurun_out/n_sanitizer_memcheck_alt.log
B200_NTT_GENERATION=2 timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/n_sanitizer_racecheck_ntt2.log 2>&1; tail -3 gpurun_out/n_sanitizer_racecheck_ntt2.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-ntt"
for lm in 6 7; do echo "== reduce_log_m=$lm 2^26" >> gpurun_out/n_bench.log; B200_MSM_REDUCE_LOG_M=$lm timeout 600 $B >> gpurun_out/n_bench.log 2>&1; done
for lm in 5 6 7; do echo "== reduce_log_m=$lm 2^23" >> gpurun_out/n_bench.log; B200_MSM_REDUCE_LOG_M=$lm timeout 600 $B --log-n-msm 23 >> gpurun_out/n_bench.log 2>&1; done
for lm in 5 6; do echo "== reduce_log_m=$lm 2^20" >> gpurun_out/n_bench.log; B200_MSM_REDUCE_LOG_M=$lm timeout 600 $B --log-n-msm 20 >> gpurun_out/n_bench.log 2>&1; done
grep -E "^==|ms_per_step" gpurun_out/n_bench.log | python3 -c "
import sys,json
lab=None
for l in sys.stdin:
    if l.startswith('=='): lab=l.strip(); continue
    try:
        d=json.loads(l); print(lab, round(d['ms_per_step'],2), {k:round(v,1) for k,v in d['phases_ms'].items()}, 'c', d['window_c'], d.get('verified'))
    except Exception as e: print(lab,'ERR',l[:200])
"
