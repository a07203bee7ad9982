#!/usr/bin/env python3
"""Fuzz sweep: many seeds of tools/fuzz_gpu.py, printing only what is NOT one of the documented deviations
(DESIGN.md section 7): PCM off the oracle, or a status pair other than (engine 0, oracle 13) / (engine 10, *).

    python tools/fuzz_sweep.py <first seed> <n seeds> [configs, default 2,3,4,5] [ds]
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fuzz_gpu

first, n = int(sys.argv[1]), int(sys.argv[2])
cfgs = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "2,3,4,5").split(",")]
ds = len(sys.argv) > 4 and sys.argv[4] == "ds"
tot = dict(frames=0, mutated=0, documented=0, bad=0)
for cfg in cfgs:
    for seed in range(first, first + n):
        r = fuzz_gpu.run(cfg, 48, 32, seed, 0.3, verbose=False, downsampled=ds and cfg in (3, 4), tile=seed % 4)
        odd = [x for x in r["bad_status"] if not ((x[2], x[3]) == (0, 13) or x[2] == 10)]
        tot["frames"] += r["frames"]; tot["mutated"] += r["mutated"]; tot["documented"] += len(r["bad_status"]) - len(odd)
        if odd or r["bad_pcm"]:
            tot["bad"] += len(odd) + len(r["bad_pcm"])
            print("config %d seed %d: status %s pcm %s" % (cfg, seed, odd, r["bad_pcm"]), flush=True)
print(tot)
