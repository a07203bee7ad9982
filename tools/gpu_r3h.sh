#!/bin/bash
# session 3, call H: JAAD's index errors inside the SBR / PS tools reported as JAADB_ST_ARRAY_BOUNDS (K3), the R channel's
# setCommonData gated on ics_info lying inside the frame; whole GPU suite, then sweeps
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > $O/r3h_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r3h_pytest.log | cut -c1-400
timeout 900 python - > $O/r3h_sweep.log 2>&1 <<'P'
import sys
sys.path.insert(0, "tools")
import fuzz_gpu
for name, cfgno, over, ds in (("c5 plain", 5, {}, False), ("c5 pulse+drc", 5, dict(p_pulse=0.6, p_drc=0.8), False), ("c3", 3, {}, False), ("c4", 4, {}, False), ("c4 ds", 4, {}, True)):
    tot = dict(frames=0, mutated=0, bad_status=0, bad_pcm=0)
    for seed in range(700, 712):
        r = fuzz_gpu.run(cfgno, 48, 32, seed, 0.35, verbose=False, gen_over=over, downsampled=ds, tile=seed % 4)
        tot["frames"] += r["frames"]; tot["mutated"] += r["mutated"]; tot["bad_status"] += len(r["bad_status"]); tot["bad_pcm"] += len(r["bad_pcm"])
        if r["bad_status"] or r["bad_pcm"]:
            print("%s seed %d: status %s pcm %s" % (name, seed, r["bad_status"], r["bad_pcm"]), flush=True)
    print(name, tot, flush=True)
P
echo "rc=$?"; tail -50 $O/r3h_sweep.log | cut -c1-300
