#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 tools/scatter_bench > gpurun_out/y_scatter_bench.jsonl 2>&1; cat gpurun_out/y_scatter_bench.jsonl
