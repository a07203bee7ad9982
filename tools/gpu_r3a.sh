#!/bin/bash
# session 3, call A: pulse_data parity (JAAD mode + JAADB_FLAG_PULSE_ISO), then the whole GPU suite, then config 2 without e2e
# (K1 took a parameter and a post-pass: the step must not have moved).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_parity_lc_gpu.py -x -q -m gpu -k pulse > $O/r3a_pulse.log 2>&1; echo "pulse rc=$?"; tail -5 $O/r3a_pulse.log | cut -c1-400
timeout 900 python -m pytest tests -x -q -m gpu > $O/r3a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r3a_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-extras --no-cpu-baseline > $O/r3a_bench_c2.json 2> $O/r3a_bench_c2.err; echo "bench rc=$?"; cut -c1-400 $O/r3a_bench_c2.json
