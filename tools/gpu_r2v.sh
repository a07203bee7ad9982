#!/bin/bash
# round-2 GPU session V: K4 part count with the aligned 10-warp K4b
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
run() { # name config env...
  local name=$1 c=$2; shift 2
  env "$@" timeout 300 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2v_bench_c${c}_$name.log 2>&1; echo c$c $name $(grep -o '"kernel_ms": {[^}]*}' $O/r2v_bench_c${c}_$name.log) $(grep -o '"bad_frames": [0-9]*' $O/r2v_bench_c${c}_$name.log | head -1)
}
for c in 3 4; do
  for p in 2 3 4; do run parts$p $c JAADB200_K4_PARTS=$p; done
done
