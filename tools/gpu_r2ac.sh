#!/bin/bash
# round-2 GPU session AC: current counters of K4a and K4c (config 3) and of K4c's two-bank PS variant (config 4)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k4a_analysis|k4c_synthesis" -s 2 -c 2 -o $O/r2ac_k4a_k4c_c3 -f python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $O/r2ac_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k4c_synthesis" -s 2 -c 1 -o $O/r2ac_k4c_c4 -f python bench.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $O/r2ac_ncu_c4.log 2>&1
