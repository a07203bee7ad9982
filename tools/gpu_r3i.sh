#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/r3i_debug.log
timeout 300 python tools/fuzz_debug.py 5 48 32 707 0.35 9 >> $O/r3i_debug.log 2>&1
cut -c1-1200 $O/r3i_debug.log | head -12
timeout 600 python -m pytest tests/test_fuzz_gpu.py -q -m gpu 2>&1 | tail -5 | cut -c1-600
