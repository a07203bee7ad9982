#!/bin/bash
# round-2 GPU session P: K5 without the energy plane (24.7 KB of shared memory, 8 CTAs per SM), 16 GB K4 tiles
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > $O/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2p_pytest.log
tail -4 $O/r2p_pytest.log
run() { # name config env...
  local name=$1 c=$2; shift 2
  env "$@" python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2p_bench_c${c}_$name.log 2>&1; echo c$c $name $(grep -o '"kernel_ms": {[^}]*}' $O/r2p_bench_c${c}_$name.log)
}
run base 4 A=1
for v in k5_mb7 k5_mb6; do run $v 4 JAADB200_LIB=jaadec_b200/_build/variants/$v.so; done
run base 3 A=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k5_ps" -s 3 -c 1 -o $O/r2p_k5_c4 -f python bench.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2p_ncu_k5.log 2>&1
