#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err; tail -c 1500 gpurun_out/w_bench.json
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/w_pytest_all.log 2>&1; tail -3 gpurun_out/w_pytest_all.log
