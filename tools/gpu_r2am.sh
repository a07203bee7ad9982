#!/bin/bash
# round-2 GPU session AM: details of corrupted-stream findings (status pairs of config 2, PCM of configs 4 / 5)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
ALLTAPS=1 timeout 300 python tools/fuzz_debug.py 2 48 32 504 0.3 19 > $O/r2am_c2_504_19.log 2>&1; tail -4 $O/r2am_c2_504_19.log | cut -c1-400
ALLTAPS=1 timeout 300 python tools/fuzz_debug.py 2 48 32 514 0.3 4 > $O/r2am_c2_514_4.log 2>&1; sed -n 10,14p $O/r2am_c2_514_4.log | cut -c1-400
PSTAPS=1 SBRTAPS=1 timeout 300 python tools/fuzz_debug.py 4 48 32 500 0.3 8 > $O/r2am_c4_500_8.log 2>&1; sed -n 26,32p $O/r2am_c4_500_8.log | cut -c1-1500
timeout 300 python tools/fuzz_debug.py 5 48 32 514 0.3 18 > $O/r2am_c5_514_18.log 2>&1; sed -n 17,22p $O/r2am_c5_514_18.log | cut -c1-900
