#!/bin/bash
# round-2 GPU session B: immediate named barriers (occupancy), warp-cooperative pre-pass, PS IPD/OPD
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2b_pytest.log
tail -5 $O/r2b_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2b_bench_mb5.log 2>&1; tail -c 700 $O/r2b_bench_mb5.log
for v in mb4 mb6; do
  JAADB200_LIB=$PWD/jaadec_b200/_build/variants/$v.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2b_bench_$v.log 2>&1; tail -c 700 $O/r2b_bench_$v.log
done
python bench.py --config 1 --steps 5 --warmup 3 --no-cpu-baseline > $O/r2b_bench_c1.log 2>&1; tail -c 900 $O/r2b_bench_c1.log
python bench.py --config 3 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > $O/r2b_bench_c3.log 2>&1; tail -c 700 $O/r2b_bench_c3.log
python bench.py --config 4 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > $O/r2b_bench_c4.log 2>&1; tail -c 700 $O/r2b_bench_c4.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2b_launches_c1.csv python bench.py --config 1 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2b_ncu0.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_filterbank -s 1 -c 1 -o $O/r2b_k2_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2b_ncu2.log 2>&1
ls -la $O | tail -12
