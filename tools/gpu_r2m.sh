#!/bin/bash
# round-2 GPU session M: K3 L1 hints / carveout variants; K4 part count against K4b's wave size
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_parity_sbr_gpu.py tests/test_fuzz_gpu.py -m gpu -q -x --timeout 900 > $O/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2m_pytest.log
tail -3 $O/r2m_pytest.log
run() { # name config env...
  local name=$1 c=$2; shift 2
  env "$@" python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2m_bench_c${c}_$name.log 2>&1; echo c$c $name $(grep -o '"kernel_ms": {[^}]*}' $O/r2m_bench_c${c}_$name.log)
}
for c in 3 4; do
  run base $c A=1
  for v in k3_nohint k3_c86 k3_c72; do run $v $c JAADB200_LIB=jaadec_b200/_build/variants/$v.so; done
  for p in 1 2 3 4 6; do run parts$p $c JAADB200_K4_PARTS=$p; done
done
