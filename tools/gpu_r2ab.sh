#!/bin/bash
# round-2 GPU session AB: six two-channel K2 CTAs per SM (80 registers, 37.8 KB of shared memory)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > $O/r2ab_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2ab_pytest.log
tail -3 $O/r2ab_pytest.log
run() { # name config env...
  local name=$1 c=$2; shift 2
  env "$@" timeout 400 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/r2ab_bench_c${c}_$name.log 2>&1; echo c$c $name $(grep -o '"kernel_ms": {[^}]*}' $O/r2ab_bench_c${c}_$name.log | head -1) $(grep -o '"bad_frames": [0-9]*' $O/r2ab_bench_c${c}_$name.log | head -1)
}
run base 2 A=1
run k2mb5 2 JAADB200_LIB=jaadec_b200/_build/variants/k2mb5.so
run base 3 A=1
run k2mb5 3 JAADB200_LIB=jaadec_b200/_build/variants/k2mb5.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k2_filterbank" -c 1 -o $O/r2ab_k2_c2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $O/r2ab_ncu.log 2>&1
