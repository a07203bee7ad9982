#!/bin/bash
# round-2 GPU session L: K3 with one expansion of every syntax function (step lists)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > $O/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2l_pytest.log
tail -4 $O/r2l_pytest.log
for c in 3 4; do
  python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2l_bench_c$c.log 2>&1; echo c$c $(grep -o '"kernel_ms": {[^}]*}' $O/r2l_bench_c$c.log)
done
for c in 3 4; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k3_sbr" -c 1 -o $O/r2l_k3_c$c -f python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2l_ncu_c$c.log 2>&1
done
ls -la $O | grep r2l | head -30
