#!/bin/bash
# round-2 GPU session AF: e2e with the PCM offsets in pinned memory and the frame-table copy next to the decode
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_demux_gpu.py tests/test_parity_lc_gpu.py -m gpu -q -x --timeout 900 > $O/r2af_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2af_pytest.log
tail -3 $O/r2af_pytest.log
JAADB200_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras > $O/r2af_bench.json 2> $O/r2af_trace.log; echo rc=$?
grep -o '"e2e": {[^}]*}' $O/r2af_bench.json | cut -c1-300
grep -n "containers indexed\|chunk 0:\|downloads done" $O/r2af_trace.log | sed -n 4,12p
