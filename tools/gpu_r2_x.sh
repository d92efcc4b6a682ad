#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 tools/scatter_bench > gpurun_out/x_scatter_bench.jsonl 2>&1; cat gpurun_out/x_scatter_bench.jsonl
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/x_smoke.log 2>&1; tail -2 gpurun_out/x_smoke.log
