#!/bin/bash
# round-2 GPU session G: host path after the indexer / fused-prep changes: tests, e2e timeline (trace), full bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > $O/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2g_pytest.log
tail -4 $O/r2g_pytest.log
JAADB200_TRACE=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/r2g_trace.log 2>&1; grep -c jaadb $O/r2g_trace.log; grep "jaadb" $O/r2g_trace.log | tail -45
( time python bench.py ) > $O/r2g_bench_full.log 2>&1; tail -c 300 $O/r2g_bench_full.log
nproc; free -g | head -2
