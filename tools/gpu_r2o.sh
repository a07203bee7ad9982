#!/bin/bash
# round-2 GPU session O: K4 tile workspace size
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
run() { # name config env...
  local name=$1 c=$2; shift 2
  env "$@" python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2o_bench_c${c}_$name.log 2>&1; echo c$c $name $(grep -o '"kernel_ms": {[^}]*}' $O/r2o_bench_c${c}_$name.log)
}
for c in 3 4; do
  run base $c A=1
  for v in tile4 tile16 tile32; do run $v $c JAADB200_LIB=jaadec_b200/_build/variants/$v.so; done
done
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
