#!/bin/bash
# session 3, call B: fill-element (dynamic range info) parity, the pulse + corrupted-stream tests on the out-of-line K1 side
# paths, then config 2 A/B: the build before pulse / DRC support (variant pre_pulse.so) against the current one, twice each.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_parity_fill_gpu.py tests/test_parity_lc_gpu.py tests/test_fuzz_gpu.py -x -q -m gpu > $O/r3b_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r3b_pytest.log | cut -c1-400
for i in 1 2; do
  JAADB200_LIB=jaadec_b200/_build/variants/pre_pulse.so timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-extras --no-cpu-baseline > $O/r3b_bench_pre_$i.json 2> $O/r3b_bench_pre_$i.err; echo "pre $i rc=$?"
  timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-extras --no-cpu-baseline > $O/r3b_bench_now_$i.json 2> $O/r3b_bench_now_$i.err; echo "now $i rc=$?"
done
python - <<'P'
import json
for n in ("pre_1","now_1","pre_2","now_2"):
    d=json.load(open("gpurun_out/r3b_bench_%s.json"%n)); print(n, d["ms_per_step"], d["roofline"]["kernel_ms"])
P
