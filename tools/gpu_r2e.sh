#!/bin/bash
# round-2 GPU session E: K2 with TMA-staged frame records + pipelined window loads; the new bench.py (all configs, demux in e2e)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > $O/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2e_pytest.log
tail -4 $O/r2e_pytest.log
JAADB200_LIB=$PWD/jaadec_b200/_build/variants/mb4.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/r2e_bench_mb4.log 2>&1
echo mb4 $(grep -o '"kernel_ms": {[^}]*}' $O/r2e_bench_mb4.log)
( time python bench.py ) > $O/r2e_bench_full.log 2>&1; tail -c 6000 $O/r2e_bench_full.log
( time python bench.py --impl reference --steps 2 --warmup 1 ) > $O/r2e_bench_ref.log 2>&1; tail -c 1500 $O/r2e_bench_ref.log
