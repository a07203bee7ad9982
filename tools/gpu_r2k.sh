#!/bin/bash
# round-2 GPU session K: K3 on staged payload words + warp-cooperative PS decode; K4b at 11 KB / warp (5 CTAs per SM)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > $O/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2k_pytest.log
tail -4 $O/r2k_pytest.log
for c in 3 4; do
  python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2k_bench_c$c.log 2>&1; echo c$c $(grep -o '"kernel_ms": {[^}]*}' $O/r2k_bench_c$c.log)
  for v in k4b_mb4 k4b_w3mb6; do
    JAADB200_LIB=jaadec_b200/_build/variants/$v.so python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2k_bench_c${c}_$v.log 2>&1; echo c$c $v $(grep -o '"kernel_ms": {[^}]*}' $O/r2k_bench_c${c}_$v.log)
  done
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2k_launches_c4.csv python bench.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2k_ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k3_sbr" -c 1 -o $O/r2k_k3_c4 -f python bench.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2k_ncu5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k4b_hf" -s 9 -c 1 -o $O/r2k_k4b_c4 -f python bench.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2k_ncu6.log 2>&1
ls -la $O | grep r2k | head -30
