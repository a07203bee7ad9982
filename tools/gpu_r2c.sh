#!/bin/bash
# round-2 GPU session C: K2 with 32-byte frame records, coalesced twiddle gather, window prefetch; occupancy x carve-out sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > $O/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c_pytest.log
tail -4 $O/r2c_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2c_bench_default.log 2>&1; grep -o '"kernel_ms": {[^}]*}' $O/r2c_bench_default.log
for v in mb4cD mb4c100 mb4c70 mb5cD mb5c80 mb6c100 mb4cDnp mb5c100np; do
  JAADB200_LIB=$PWD/jaadec_b200/_build/variants/$v.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2c_bench_$v.log 2>&1
  echo $v $(grep -o '"kernel_ms": {[^}]*}' $O/r2c_bench_$v.log)
done
python bench.py --config 5 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --streams 4096 > $O/r2c_bench_c5.log 2>&1; grep -o '"kernel_ms": {[^}]*}' $O/r2c_bench_c5.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_filterbank -s 1 -c 1 -o $O/r2c_k2_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2c_ncu2.log 2>&1
JAADB200_LIB=$PWD/jaadec_b200/_build/variants/mb4cD.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_filterbank -s 1 -c 1 -o $O/r2c_k2_full_mb4cD -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2c_ncu3.log 2>&1
ls -la $O | tail -6
