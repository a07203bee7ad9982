#!/bin/bash
# round-2 GPU session I: K5 without staged matrices (6 CTAs / SM), K3 with 8-bit Huffman tables; host prep on threads
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > $O/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2i_pytest.log
tail -4 $O/r2i_pytest.log
for c in 3 4; do
  python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2i_bench_c$c.log 2>&1; echo c$c $(grep -o '"kernel_ms": {[^}]*}' $O/r2i_bench_c$c.log)
done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/r2i_bench_c2.log 2>&1; grep -o '"e2e": {[^}]*}' $O/r2i_bench_c2.log | cut -c1-200; grep -o '"e2e_device": {[^}]*}' $O/r2i_bench_c2.log | cut -c1-120
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2i_launches_c4.csv python bench.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2i_ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k5_ps|k3_sbr|k4b_hf" -s 9 -c 3 -o $O/r2i_k3_k4b_k5_full -f python bench.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2i_ncu5.log 2>&1
