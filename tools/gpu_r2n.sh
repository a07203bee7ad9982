#!/bin/bash
# round-2 GPU session N: batches with SBR streams go down K3 / K2 / K4 in parts (K3 + K2 of part p+1 under part p's QMF pipeline)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > $O/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2n_pytest.log
tail -4 $O/r2n_pytest.log
run() { # name config env...
  local name=$1 c=$2; shift 2
  env "$@" python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2n_bench_c${c}_$name.log 2>&1; echo c$c $name $(grep -o '"kernel_ms": {[^}]*}' $O/r2n_bench_c${c}_$name.log)
}
for c in 3 4; do
  for p in 1 2 3 4 6; do run parts$p $c JAADB200_K4_PARTS=$p; done
done
run base 2 A=1
