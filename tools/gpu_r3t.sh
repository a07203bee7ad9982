#!/bin/bash
# session 3, call T: config 2 A/B on one box -- the build before K1's frames-per-warp mapping against the one with it
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for i in 1 2; do
  JAADB200_LIB=jaadec_b200/_build/variants/before_lanes.so timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-extras --no-cpu-baseline > $O/r3t_pre_$i.json 2> $O/r3t_pre_$i.err; echo "pre $i rc=$?"
  timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-extras --no-cpu-baseline > $O/r3t_now_$i.json 2> $O/r3t_now_$i.err; echo "now $i rc=$?"
done
python - <<'P'
import json
for n in ("pre_1","now_1","pre_2","now_2"):
    d=json.load(open("gpurun_out/r3t_%s.json"%n)); print(n, d["ms_per_step"], d["roofline"]["kernel_ms"])
P
