#!/bin/bash
# round-2 GPU session AE: host-side timeline of the e2e call (JAADB200_TRACE)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
JAADB200_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-extras > $O/r2ae_bench.json 2> $O/r2ae_trace.log; echo rc=$?
grep -c . $O/r2ae_trace.log; tail -60 $O/r2ae_trace.log | cut -c1-200
