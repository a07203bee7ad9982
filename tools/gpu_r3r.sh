#!/bin/bash
# session 3, call R: the corrupted-stream sweep of session AL (same seeds, same mutations) on the final build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python tools/fuzz_sweep.py 500 24 2,3,4,5 > $O/r3r_fuzz.log 2>&1; echo "sweep rc=$?"; tail -14 $O/r3r_fuzz.log | cut -c1-300
timeout 900 python tools/fuzz_sweep.py 700 12 3,4 ds > $O/r3r_fuzz_ds.log 2>&1; echo "sweep ds rc=$?"; tail -6 $O/r3r_fuzz_ds.log | cut -c1-300
