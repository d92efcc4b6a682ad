#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in "" _g2v3 _g2v4 _g2v5; do echo "######## lib$v" >> gpurun_out/e_g2_debug.log; B200_LIB_PATH=$PWD/algebra_b200/libalgebra_b200$v.so timeout 600 python tools/g2_debug.py 2>&1 | grep -A7 "===== G2" >> gpurun_out/e_g2_debug.log; done
cat gpurun_out/e_g2_debug.log
for g in 1 2; do for ln in 20 22 24 26; do B200_NTT_GENERATION=$g timeout 300 python tools/ntt_time.py --log-n $ln >> gpurun_out/e_ntt_time.jsonl 2>&1; done; done
grep -v roundtrip gpurun_out/e_ntt_time.jsonl
timeout 900 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_primitives.py -x -q > gpurun_out/e_pytest.log 2>&1; tail -3 gpurun_out/e_pytest.log
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-ntt"
run() { name=$1; shift; echo "== $name" >> gpurun_out/e_variants.log; env "$@" timeout 600 $B >> gpurun_out/e_variants.log 2>&1; }
run default X=1
run cap5_load8 B200_MSM_LEVEL_CAP=5 B200_MSM_LEVEL_MIN_LOAD=8
run cap6_load4 B200_MSM_LEVEL_CAP=6 B200_MSM_LEVEL_MIN_LOAD=4
run cap6_load4_min64 B200_MSM_LEVEL_CAP=6 B200_MSM_LEVEL_MIN_LOAD=4 B200_MSM_MIN_BATCH=64
grep -E "^==|ms_per_step" gpurun_out/e_variants.log | cut -c1-330
for ln in 22 23 24 25; do for c in 15 16 17 18 19; do
  echo "== n=2^$ln c=$c" >> gpurun_out/e_sweep.log
  B200_MSM_LEVEL_CAP=6 B200_MSM_LEVEL_MIN_LOAD=4 timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-ntt --no-verify --log-n-msm $ln --window $c >> gpurun_out/e_sweep.log 2>&1
done; done
for c in 19 21; do echo "== n=2^26 c=$c" >> gpurun_out/e_sweep.log; B200_MSM_LEVEL_CAP=6 B200_MSM_LEVEL_MIN_LOAD=4 timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-ntt --no-verify --window $c >> gpurun_out/e_sweep.log 2>&1; done
grep -E "^==|ms_per_step" gpurun_out/e_sweep.log | cut -c1-200
