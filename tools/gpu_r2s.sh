#!/bin/bash
# round-2 GPU session S: K4b warps of a CTA aligned per phase (instruction caches), CTA sizes 4 / 10 / 20 warps
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
run() { # name config env...
  local name=$1 c=$2; shift 2
  env "$@" python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2s_bench_c${c}_$name.log 2>&1; echo c$c $name $(grep -o '"kernel_ms": {[^}]*}' $O/r2s_bench_c${c}_$name.log) $(grep -o '"bad_frames": [0-9]*' $O/r2s_bench_c${c}_$name.log | head -1)
}
for c in 3 4; do
  run base $c A=1
  for v in w4al w20 w20al w10al; do run $v $c JAADB200_LIB=jaadec_b200/_build/variants/$v.so; done
done
JAADB200_LIB=jaadec_b200/_build/variants/w20al.so timeout 900 python -m pytest tests/test_parity_sbr_gpu.py -m gpu -q -x --timeout 900 > $O/r2s_pytest.log 2>&1; echo "pytest w20al rc=$?" >> $O/r2s_pytest.log
tail -3 $O/r2s_pytest.log
