#!/bin/bash
# session 3, call G: window-shape notes for elements outside the layout that name the stream's own objects (5.1 id / tag flips)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_fuzz_gpu.py tests/test_parity_lc_gpu.py -q -m gpu > $O/r3g_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r3g_pytest.log | cut -c1-400
timeout 900 python - > $O/r3g_sweep.log 2>&1 <<'P'
import sys
sys.path.insert(0, "tools")
import fuzz_gpu
for name, cfgno, over in (("c5 plain", 5, {}), ("c5 pulse+drc", 5, dict(p_pulse=0.6, p_drc=0.8)), ("c2 plain", 2, {}), ("c4", 4, {})):
    tot = dict(frames=0, mutated=0, bad_status=0, bad_pcm=0)
    for seed in range(700, 712):
        r = fuzz_gpu.run(cfgno, 48, 32, seed, 0.35, verbose=False, gen_over=over)
        tot["frames"] += r["frames"]; tot["mutated"] += r["mutated"]; tot["bad_status"] += len(r["bad_status"]); tot["bad_pcm"] += len(r["bad_pcm"])
        if r["bad_status"] or r["bad_pcm"]:
            print("%s seed %d: status %s pcm %s" % (name, seed, r["bad_status"], r["bad_pcm"]), flush=True)
    print(name, tot, flush=True)
P
echo "rc=$?"; tail -40 $O/r3g_sweep.log | cut -c1-300
