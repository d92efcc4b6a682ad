#!/bin/bash
# round-2 GPU batch G (8 GPUs): multi-rank parity (NCCL world 2, single-process multi-device C ABI), strong-scaling bench at N = 8, 4, 2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/g_gpus.txt
timeout 900 python -m pytest tests/test_gpu_dist_nccl.py tests/test_cpp_mirror.py -q > gpurun_out/g_pytest_dist.log 2>&1; echo "rc=$?" >> gpurun_out/g_pytest_dist.log; tail -4 gpurun_out/g_pytest_dist.log
P=29500
for n in 8 4 2; do
  P=$((P+1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_${n}gpu.json 2> gpurun_out/g_bench_${n}gpu.err
  tail -c 1200 gpurun_out/g_bench_${n}gpu.json; echo
done
# the same MSM through the single-process C ABI on 1..8 devices (host buffers, pageable memory)
timeout 900 python tools/multi_capi_bench.py > gpurun_out/g_multi_capi.jsonl 2>&1; cat gpurun_out/g_multi_capi.jsonl
