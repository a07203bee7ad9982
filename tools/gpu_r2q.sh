#!/bin/bash
# round-2 GPU session Q: full default bench line (+ extras) of the build with the K3 / K4b / K5 changes, launch lists
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > $O/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2q_pytest.log
tail -3 $O/r2q_pytest.log
python bench.py > $O/r2q_bench_default.json 2> $O/r2q_bench_default.err; echo "bench rc=$?"; cut -c1-1500 $O/r2q_bench_default.json
for c in 2 3 4; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2q_launches_c$c.csv python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $O/r2q_ncu_c$c.log 2>&1
done
