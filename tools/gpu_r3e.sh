#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python - > $O/r3e_sweep.log 2>&1 <<'P'
import sys
sys.path.insert(0, "tools")
import fuzz_gpu
for name, over in (("plain", {}), ("drc", dict(p_drc=0.8)), ("pulse", dict(p_pulse=0.6))):
    tot = dict(frames=0, mutated=0, bad_status=0, bad_pcm=0)
    for seed in range(700, 712):
        r = fuzz_gpu.run(5, 48, 32, seed, 0.35, verbose=False, gen_over=over)
        tot["frames"] += r["frames"]; tot["mutated"] += r["mutated"]; tot["bad_status"] += len(r["bad_status"]); tot["bad_pcm"] += len(r["bad_pcm"])
        if r["bad_status"] or r["bad_pcm"]:
            print("%s seed %d: status %s pcm %s" % (name, seed, r["bad_status"], r["bad_pcm"]), flush=True)
    print(name, tot, flush=True)
P
echo "rc=$?"; tail -40 $O/r3e_sweep.log | cut -c1-300
