#!/bin/bash
# round-2 GPU session W: K4b L2 prefetch placement (bit 0: generated band before the gains, bit 1: next frame's low band before the assembly)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
run() { # name config env...
  local name=$1 c=$2; shift 2
  env "$@" timeout 300 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2w_bench_c${c}_$name.log 2>&1; echo c$c $name $(grep -o '"kernel_ms": {[^}]*}' $O/r2w_bench_c${c}_$name.log) $(grep -o '"bad_frames": [0-9]*' $O/r2w_bench_c${c}_$name.log | head -1)
}
for c in 3 4; do
  run base $c A=1
  for v in pf0 pf1 pf2; do run $v $c JAADB200_LIB=jaadec_b200/_build/variants/$v.so; done
done
timeout 900 python -m pytest tests/test_parity_sbr_gpu.py tests/test_fuzz_gpu.py -m gpu -q -x --timeout 900 > $O/r2w_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2w_pytest.log
tail -3 $O/r2w_pytest.log
