#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for p in 2 4 8; do timeout 300 python tools/overlap_probe.py 4096 $p >> $O/r3c_overlap.log 2>&1; echo "rc=$?"; done
cat $O/r3c_overlap.log | tail -5
