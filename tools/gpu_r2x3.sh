#!/bin/bash
# round-2 GPU session X3: validation of the final build (K4b with two alignment barriers per frame): GPU suite, smoke, default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > $O/r2x3_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2x3_pytest.log
tail -3 $O/r2x3_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2x3_smoke.log 2>&1; tail -2 $O/r2x3_smoke.log
python bench.py > $O/r2x3_bench_default.json 2> $O/r2x3_bench_default.err; echo "bench rc=$?"; cut -c1-200 $O/r2x3_bench_default.json
for c in 3 4; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2x3_launches_c$c.csv python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $O/r2x3_ncu_c$c.log 2>&1
done
