#!/bin/bash
# session 3, call N: the final build -- whole GPU suite, smoke, default bench line, reference arm, launch lists of configs 3 / 4
# (K3 changed after call K; K1 / K2 did not: their launch list and --set full counters of call K stand)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > $O/r3n_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r3n_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r3n_smoke.log 2>&1; tail -2 $O/r3n_smoke.log
python bench.py > $O/r3n_bench_default.json 2> $O/r3n_bench_default.err; echo "bench rc=$?"; cut -c1-300 $O/r3n_bench_default.json
python bench.py --impl reference > $O/r3n_bench_reference.json 2> $O/r3n_bench_reference.err; echo "ref rc=$?"; cut -c1-200 $O/r3n_bench_reference.json
for c in 3 4; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r3n_launches_c$c.csv python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $O/r3n_ncu_c$c.log 2>&1; echo "launch list $c rc=$?"
done
