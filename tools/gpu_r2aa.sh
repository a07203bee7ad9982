#!/bin/bash
# round-2 GPU session AA: three- to six-channel streams on a 384-thread K2 instantiation with two CTAs per SM (80 registers)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > $O/r2aa_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2aa_pytest.log
tail -3 $O/r2aa_pytest.log
run() { # name config env...
  local name=$1 c=$2; shift 2
  env "$@" timeout 400 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2aa_bench_c${c}_$name.log 2>&1; echo c$c $name $(grep -o '"kernel_ms": {[^}]*}' $O/r2aa_bench_c${c}_$name.log) $(grep -o '"bad_frames": [0-9]*' $O/r2aa_bench_c${c}_$name.log | head -1)
}
run base 5 A=1
run mc1 5 JAADB200_LIB=jaadec_b200/_build/variants/mc1.so
timeout 600 ncu --set full --clock-control none -k regex:"k2_filterbank" -c 1 -o $O/r2aa_k2_c5 -f python bench.py --config 5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2aa_ncu.log 2>&1
