#!/bin/bash
# session 3, call K: the final build -- smoke, default bench line, reference arm, launch list of config 2, --set full
# counters of K1 / K2 (K1 gained the out-of-line side paths and the shape notes this session)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > $O/r3k_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r3k_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r3k_smoke.log 2>&1; tail -2 $O/r3k_smoke.log
python bench.py > $O/r3k_bench_default.json 2> $O/r3k_bench_default.err; echo "bench rc=$?"; cut -c1-300 $O/r3k_bench_default.json
python bench.py --impl reference > $O/r3k_bench_reference.json 2> $O/r3k_bench_reference.err; echo "ref rc=$?"; cut -c1-300 $O/r3k_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r3k_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $O/r3k_ncu_c2.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k1_parse|k2_filterbank" -s 2 -c 2 -o $O/r3k_k1_k2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $O/r3k_ncu_full.log 2>&1; echo "set full rc=$?"
python tools/ncuread.py $O/r3k_k1_k2.ncu-rep > $O/r3k_k1_k2_ncu_raw.txt 2>&1; tail -5 $O/r3k_k1_k2_ncu_raw.txt
