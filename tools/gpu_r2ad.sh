#!/bin/bash
# round-2 GPU session AD: K4a with the warps of a CTA aligned per phase
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
run() { # name config env...
  local name=$1 c=$2; shift 2
  env "$@" timeout 400 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/r2ad_bench_c${c}_$name.log 2>&1; echo c$c $name $(grep -o '"kernel_ms": {[^}]*}' $O/r2ad_bench_c${c}_$name.log | head -1) $(grep -o '"bad_frames": [0-9]*' $O/r2ad_bench_c${c}_$name.log | head -1)
}
for c in 3 4; do
  run base $c A=1
  for v in a4al a8al a10al a16al a10; do run $v $c JAADB200_LIB=jaadec_b200/_build/variants/$v.so; done
done
JAADB200_LIB=jaadec_b200/_build/variants/a10al.so timeout 900 python -m pytest tests/test_parity_sbr_gpu.py -m gpu -q -x --timeout 900 > $O/r2ad_pytest.log 2>&1; echo "pytest a10al rc=$?" >> $O/r2ad_pytest.log
tail -3 $O/r2ad_pytest.log
