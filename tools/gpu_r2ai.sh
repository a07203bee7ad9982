#!/bin/bash
# round-2 GPU session AI: K5's IPD/OPD phase bookkeeping on one thread per parameter band
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > $O/r2ai_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2ai_pytest.log
tail -3 $O/r2ai_pytest.log
for i in 1 2; do
timeout 400 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/r2ai_bench_c4_$i.log 2>&1; echo c4 $(grep -o '"kernel_ms": {[^}]*}' $O/r2ai_bench_c4_$i.log | head -1) $(grep -o '"bad_frames": [0-9]*' $O/r2ai_bench_c4_$i.log | head -1)
done
