#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist_nccl.py tests/test_cpp_mirror.py -x -q -m gpu > gpurun_out/v_pytest_dist.log 2>&1; tail -3 gpurun_out/v_pytest_dist.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3"
timeout 900 $T > gpurun_out/v_bench_2gpu_buckets.json 2> gpurun_out/v_bench_2gpu_buckets.err; tail -c 3000 gpurun_out/v_bench_2gpu_buckets.json
timeout 900 $T --shard chunks --no-ntt > gpurun_out/v_bench_2gpu_chunks.json 2> gpurun_out/v_bench_2gpu_chunks.err; tail -c 1500 gpurun_out/v_bench_2gpu_chunks.json
