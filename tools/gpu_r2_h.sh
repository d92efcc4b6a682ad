#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for f in 1 0; do for ln in 20 24 26; do B200_NTT_FUSE_LAST2=$f timeout 300 python tools/ntt_time.py --log-n $ln 2>&1 | grep -v roundtrip | sed "s/^/fuse=$f /" >> gpurun_out/h_ntt_time.log; done; done
cat gpurun_out/h_ntt_time.log
timeout 900 python -m pytest tests/test_gpu_ntt.py tests/test_cpp_mirror.py -x -q > gpurun_out/h_pytest.log 2>&1; tail -3 gpurun_out/h_pytest.log
for rep in 1 2; do
 for w in 0 17 16; do echo "== 2^24 window=$w rep=$rep" >> gpurun_out/h_2e24.log; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-ntt --no-verify --log-n-msm 24 --window $w >> gpurun_out/h_2e24.log 2>&1; done
done
grep -E "^==|ms_per_step" gpurun_out/h_2e24.log | cut -c1-260
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-ntt > gpurun_out/h_bench_e2e.json 2>&1; tail -c 900 gpurun_out/h_bench_e2e.json
