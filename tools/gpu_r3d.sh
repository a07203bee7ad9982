#!/bin/bash
# session 3, call D: corrupted streams with dynamic-range-info fill elements and pulse data (the new out-of-line K1 paths)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_fuzz_gpu.py -q -m gpu > $O/r3d_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r3d_pytest.log | cut -c1-600
timeout 900 python - > $O/r3d_sweep.log 2>&1 <<'P'
import sys
sys.path.insert(0, "tools")
import fuzz_gpu
tot = dict(frames=0, mutated=0, bad=0)
for cfg, over, iso in ((2, dict(p_drc=0.8, p_pulse=0.6), False), (2, dict(p_drc=0.8, p_pulse=0.8, pulse_wild=True), True),
                       (5, dict(p_drc=0.8, p_pulse=0.6), False), (1, dict(p_drc=0.8, p_pulse=0.9), True)):
    for seed in range(700, 710):
        r = fuzz_gpu.run(cfg, 48, 32, seed, 0.35, verbose=False, gen_over=over, pulse_iso=iso)
        tot["frames"] += r["frames"]; tot["mutated"] += r["mutated"]
        if r["bad_status"] or r["bad_pcm"]:
            tot["bad"] += len(r["bad_status"]) + len(r["bad_pcm"])
            print("config %d iso %d seed %d: status %s pcm %s" % (cfg, iso, seed, r["bad_status"], r["bad_pcm"]), flush=True)
        print("config %d iso %d seed %d statuses %s unsupported %d foreign %d" % (cfg, iso, seed, dict(sorted(r["oracle_statuses"].items())), r["unsupported"], r["foreign"]), flush=True)
print(tot)
P
echo "sweep rc=$?"; grep -v "statuses" $O/r3d_sweep.log | tail -12 | cut -c1-400
