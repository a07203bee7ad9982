#!/bin/bash
# round-2 GPU session AN: the corrupted-stream sweep of session AL against the engine as it was before this session's K3 / K4b / K5
# rewrites (commit fd3fef8 built as a variant): are the findings the documented classes, or new?
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
JAADB200_LIB=jaadec_b200/_build/variants/pre.so timeout 1500 python tools/fuzz_sweep.py 500 24 3,4,5 > $O/r2an_fuzz_pre.log 2>&1; echo "sweep pre rc=$?"; tail -3 $O/r2an_fuzz_pre.log | cut -c1-300
timeout 1500 python tools/fuzz_sweep.py 500 24 3,4,5 > $O/r2an_fuzz_now.log 2>&1; echo "sweep now rc=$?"; tail -3 $O/r2an_fuzz_now.log | cut -c1-300
diff <(grep "^config" $O/r2an_fuzz_pre.log) <(grep "^config" $O/r2an_fuzz_now.log) > $O/r2an_diff.log; echo "diff rc=$?"; head -20 $O/r2an_diff.log | cut -c1-300
