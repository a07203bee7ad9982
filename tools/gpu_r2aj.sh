#!/bin/bash
# round-2 GPU session AJ: K5 without the barrier at the top of the frame loop (hybrid synthesis next to the next frame's staging)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_parity_sbr_gpu.py tests/test_parity_sbr_downsampled_gpu.py tests/test_fuzz_gpu.py tests/test_full_size_gpu.py -m gpu -q -x --timeout 900 > $O/r2aj_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2aj_pytest.log
tail -3 $O/r2aj_pytest.log
for v in base k5tb base k5tb; do
if [ $v = base ]; then L=""; else L="JAADB200_LIB=jaadec_b200/_build/variants/$v.so"; fi
env $L timeout 400 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/r2aj_bench_c4_$v.log 2>&1; echo c4 $v $(grep -o '"kernel_ms": {[^}]*}' $O/r2aj_bench_c4_$v.log | head -1) $(grep -o '"bad_frames": [0-9]*' $O/r2aj_bench_c4_$v.log | head -1)
done
