#!/bin/bash
# round-2 GPU batch F: full parity suite on the final build, the bench lines, traffic / launch lists and full ncu captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/f_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest_gpu.log; tail -6 gpurun_out/f_pytest_gpu.log | cut -c1-200
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; tail -c 1500 gpurun_out/f_bench.json
for cfg in "--log-n-msm 24" "--log-n-msm 20" "--log-n-msm 16" "--curve 1 --log-n-msm 24"; do
  echo "== $cfg" >> gpurun_out/f_bench_configs.log
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ntt $cfg >> gpurun_out/f_bench_configs.log 2>&1
done
grep -E "^==|ms_per_step" gpurun_out/f_bench_configs.log | cut -c1-300
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
timeout 900 $NCU --log-file gpurun_out/f_traffic_msm.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-ntt --no-verify > gpurun_out/f_traffic_msm.log 2>&1
timeout 600 $NCU --log-file gpurun_out/f_traffic_ntt.csv python tools/ntt_time.py --log-n 24 --reps 1 > gpurun_out/f_traffic_ntt.log 2>&1
FULL="ncu --set full --import-source on --clock-control none"
BS="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-ntt --no-verify --log-n-msm 24 --window 18"
timeout 900 $FULL -k regex:msm_pair_add2 -c 2 -o gpurun_out/f_pair2 -f $BS > gpurun_out/f_ncu_pair2.log 2>&1
timeout 600 $FULL -k regex:ntt_pass -c 3 -o gpurun_out/f_ntt1 -f python tools/ntt_time.py --log-n 24 --reps 1 > gpurun_out/f_ncu_ntt1.log 2>&1
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 --ref-full > gpurun_out/f_reference_full.json 2> gpurun_out/f_reference_full.err; cat gpurun_out/f_reference_full.json | cut -c1-900
ls -la gpurun_out | tail -15
