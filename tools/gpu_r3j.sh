#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python - > $O/r3j.log 2>&1 <<'P'
import sys
sys.path.insert(0, "tools")
import fuzz_gpu
for a in ((3, 5, 0, False), (4, 6, 3, False), (3, 15, 5, True), (4, 16, 0, True)):
    r = fuzz_gpu.run(a[0], 24, 24, a[1], 0.3, tile=a[2], verbose=False, downsampled=a[3])
    print(a, r["bad_status"], r["bad_pcm"], r["mutated"])
for seed in range(700, 712):
    r = fuzz_gpu.run(5, 48, 32, seed, 0.35, verbose=False)
    print("c5", seed, r["bad_status"], r["bad_pcm"])
P
cat $O/r3j.log | cut -c1-300
timeout 600 python -m pytest tests/test_fuzz_gpu.py tests/test_parity_lc_gpu.py -q -m gpu 2>&1 | tail -5 | cut -c1-600
