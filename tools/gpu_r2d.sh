#!/bin/bash
# round-2 GPU session D: K2 instruction-footprint variants (rolled phase 1, phase-3 unroll), device-resident output test
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_parity_lc_gpu.py -m gpu -q -x --timeout 600 > $O/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2d_pytest.log
tail -4 $O/r2d_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2d_bench_default.log 2>&1; echo default $(grep -o '"kernel_ms": {[^}]*}' $O/r2d_bench_default.log)
for v in p3u4 p3u2 p3u1 mb5p3u2 mb5; do
  JAADB200_LIB=$PWD/jaadec_b200/_build/variants/$v.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2d_bench_$v.log 2>&1
  echo $v $(grep -o '"kernel_ms": {[^}]*}' $O/r2d_bench_$v.log)
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_filterbank -s 1 -c 1 -o $O/r2d_k2_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2d_ncu2.log 2>&1
JAADB200_LIB=$PWD/jaadec_b200/_build/variants/p3u1.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_filterbank -s 1 -c 1 -o $O/r2d_k2_full_p3u1 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2d_ncu3.log 2>&1
