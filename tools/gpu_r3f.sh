#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for c in "701 7" "701 3" "704 14" "707 9" "701 25"; do set -- $c
  echo "=== seed $1 stream $2" >> $O/r3f_debug.log
  timeout 300 python tools/fuzz_debug.py 5 48 32 $1 0.35 $2 >> $O/r3f_debug.log 2>&1
done
cut -c1-900 $O/r3f_debug.log | tail -120
