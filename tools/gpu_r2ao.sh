#!/bin/bash
# round-2 GPU session AO: ncu --set full of K1, K2 pre-pass and K2 on the default workload, final build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $O/r2ao_plain.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k1_parse|k2_prepass|k2_filterbank" -c 3 -o $O/r2ao_k1_k2_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $O/r2ao_ncu.log 2>&1; echo "ncu rc=$?"
