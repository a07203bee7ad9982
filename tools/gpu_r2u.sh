#!/bin/bash
# round-2 GPU session U: profile of the aligned K4b (config 3), launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k4b_hf" -s 3 -c 1 -o $O/r2u_k4b_c3 -f python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2u_ncu_k4b.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2u_launches_c3.csv python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $O/r2u_ncu_c3.log 2>&1
