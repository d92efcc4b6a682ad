#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/i_smoke.log 2>&1; tail -2 gpurun_out/i_smoke.log
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/i_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest_gpu.log; tail -4 gpurun_out/i_pytest_gpu.log | cut -c1-200
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; tail -c 700 gpurun_out/i_bench.json
for cfg in "--curve 2 --log-n-msm 22" "--curve 2 --log-n-msm 20"; do
  echo "== $cfg" >> gpurun_out/i_bench_g2.log
  timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-ntt $cfg >> gpurun_out/i_bench_g2.log 2>&1
done
grep -E "^==|ms_per_step" gpurun_out/i_bench_g2.log | cut -c1-400
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
timeout 600 $NCU --log-file gpurun_out/i_traffic_ntt.csv python tools/ntt_time.py --log-n 24 --reps 1 > gpurun_out/i_traffic_ntt.log 2>&1
FULL="ncu --set full --import-source on --clock-control none"
timeout 600 $FULL -k regex:ntt_pass -c 3 -o gpurun_out/i_ntt1 -f python tools/ntt_time.py --log-n 24 --reps 1 > gpurun_out/i_ncu_ntt1.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/i_reference.json 2> gpurun_out/i_reference.err; cut -c1-600 gpurun_out/i_reference.json
