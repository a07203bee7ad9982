#!/bin/bash
# round-2 GPU session T: K3 with the warps of a CTA aligned per frame, CTA sizes 4 / 7 / 14 / 28 warps; K4b 5 vs 10 warps
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
run() { # name config env...
  local name=$1 c=$2; shift 2
  env "$@" timeout 300 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2t_bench_c${c}_$name.log 2>&1; echo c$c $name $(grep -o '"kernel_ms": {[^}]*}' $O/r2t_bench_c${c}_$name.log) $(grep -o '"bad_frames": [0-9]*' $O/r2t_bench_c${c}_$name.log | head -1)
}
for c in 3 4; do
  run base $c A=1
  for v in k3w4al k3w7al k3w14al k3w28al k4bw5al; do run $v $c JAADB200_LIB=jaadec_b200/_build/variants/$v.so; done
done
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > $O/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2t_pytest.log
tail -3 $O/r2t_pytest.log
