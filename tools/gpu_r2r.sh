#!/bin/bash
# round-2 GPU session R: next-frame L2 prefetch in K4b and K5
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_parity_sbr_gpu.py tests/test_parity_sbr_downsampled_gpu.py -m gpu -q -x --timeout 900 > $O/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2r_pytest.log
tail -3 $O/r2r_pytest.log
run() { # name config env...
  local name=$1 c=$2; shift 2
  env "$@" python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2r_bench_c${c}_$name.log 2>&1; echo c$c $name $(grep -o '"kernel_ms": {[^}]*}' $O/r2r_bench_c${c}_$name.log)
}
for c in 3 4; do
  run base $c A=1
  for v in k4bpf0 k5pf0 nopf; do run $v $c JAADB200_LIB=jaadec_b200/_build/variants/$v.so; done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k4b_hf" -s 3 -c 1 -o $O/r2r_k4b_c3 -f python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2r_ncu_k4b.log 2>&1
