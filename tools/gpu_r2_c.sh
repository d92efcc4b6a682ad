#!/bin/bash
# round-2 GPU batch C: gather layouts, G2 debug, hybrid pair-add variants, window sweeps at shard sizes, DRAM-traffic launch lists
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
./tools/gather_bench > gpurun_out/c_gather.jsonl 2>&1; cat gpurun_out/c_gather.jsonl
timeout 900 python -m pytest tests/test_gpu_g2.py -q > gpurun_out/c_pytest_g2.log 2>&1; echo "rc=$?" >> gpurun_out/c_pytest_g2.log; tail -30 gpurun_out/c_pytest_g2.log | cut -c1-200
timeout 1200 python -m pytest tests/test_gpu_msm.py tests/test_gpu_primitives.py tests/test_gpu_dist_nccl.py -x -q > gpurun_out/c_pytest_msm.log 2>&1; echo "rc=$?" >> gpurun_out/c_pytest_msm.log; tail -5 gpurun_out/c_pytest_msm.log
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-ntt"
run() { name=$1; shift; echo "== $name" >> gpurun_out/c_variants.log; env "$@" timeout 600 $B >> gpurun_out/c_variants.log 2>&1; }
run default_hybrid X=1
run l1_v2_ca B200_MSM_PAIR_VARIANT_L1=2
run all_v1 B200_MSM_PAIR_VARIANT=1
grep -E "^==|ms_per_step" gpurun_out/c_variants.log | cut -c1-330
# window sweeps at shard sizes (per-rank inputs of the 2/4/8-GPU runs) for the automatic window model
for ln in 23 24 25; do for c in 16 17 18 19 20 21; do
  echo "== n=2^$ln c=$c" >> gpurun_out/c_sweep.log
  timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-ntt --no-verify --log-n-msm $ln --window $c >> gpurun_out/c_sweep.log 2>&1
done; done
grep -E "^==|ms_per_step" gpurun_out/c_sweep.log | cut -c1-260
# DRAM traffic per kernel at the metric configs (roofline.traffic)
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
timeout 900 $NCU --log-file gpurun_out/c_traffic_msm.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-ntt --no-verify > gpurun_out/c_traffic_msm.log 2>&1
B200_NTT_GENERATION=1 timeout 600 $NCU --log-file gpurun_out/c_traffic_ntt.csv python tools/ntt_time.py --log-n 24 --reps 1 > gpurun_out/c_traffic_ntt.log 2>&1
ls -la gpurun_out | tail -12
