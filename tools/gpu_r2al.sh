#!/bin/bash
# round-2 GPU session AL: corrupted-stream sweep of the final build (K3 step lists / staged reader, aligned K4b, K5 row sums)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python tools/fuzz_sweep.py 500 24 2,3,4,5 > $O/r2al_fuzz.log 2>&1; echo "sweep rc=$?"; tail -12 $O/r2al_fuzz.log | cut -c1-300
timeout 900 python tools/fuzz_sweep.py 700 12 3,4 ds > $O/r2al_fuzz_ds.log 2>&1; echo "sweep ds rc=$?"; tail -6 $O/r2al_fuzz_ds.log | cut -c1-300
