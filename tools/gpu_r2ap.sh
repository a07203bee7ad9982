#!/bin/bash
# round-2 GPU session AP: K4b phase alignment with four, three or two barriers per frame
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
run() { # name config env...
  local name=$1 c=$2; shift 2
  env "$@" timeout 400 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/r2ap_bench_c${c}_$name.log 2>&1; echo c$c $name $(grep -o '"kernel_ms": {[^}]*}' $O/r2ap_bench_c${c}_$name.log | head -1) $(grep -o '"bad_frames": [0-9]*' $O/r2ap_bench_c${c}_$name.log | head -1)
}
for c in 3 4; do
  run base $c A=1
  for v in al2 al3; do run $v $c JAADB200_LIB=jaadec_b200/_build/variants/$v.so; done
  run base2 $c A=1
done
for v in al2 al3; do
JAADB200_LIB=jaadec_b200/_build/variants/$v.so timeout 900 python -m pytest tests/test_parity_sbr_gpu.py -m gpu -q -x --timeout 900 > $O/r2ap_pytest_$v.log 2>&1; echo "pytest $v rc=$?"; tail -1 $O/r2ap_pytest_$v.log
done
