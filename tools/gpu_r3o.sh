#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
python bench.py --no-extras --no-cpu-baseline > $O/r3o_bench_c2.json 2> $O/r3o_bench_c2.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open("gpurun_out/r3o_bench_c2.json")); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"]["step_device_total"], d["e2e"]["value"], d["bad_frames"], d["bad_frames_repeat_pass"])
P
