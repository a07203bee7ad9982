#!/bin/bash
# session 3, call S: K1 with fewer frames per warp for small batches -- whole GPU suite (small batches: one frame per warp),
# config 1 and config 2 (dense mapping: must not move)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > $O/r3s_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r3s_pytest.log
for l in 5 0; do
JAADB_K1_LANES_LOG2=$l timeout 300 python bench.py --config 1 --steps 20 --warmup 5 --no-e2e --no-extras --no-cpu-baseline > $O/r3s_c1_l$l.json 2> $O/r3s_c1_l$l.err; echo "c1 l=$l rc=$?"
done
timeout 300 python bench.py --config 1 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $O/r3s_c1_auto.json 2> $O/r3s_c1_auto.err; echo "c1 auto rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-extras --no-cpu-baseline > $O/r3s_c2.json 2> $O/r3s_c2.err; echo "c2 rc=$?"
python - <<'P'
import json
for n in ("c1_l5","c1_l0","c1_auto","c2"):
    d=json.load(open("gpurun_out/r3s_%s.json"%n)); print(n, round(d["value"]), d["ms_per_step"], d["roofline"]["kernel_ms"], (d.get("e2e") or {}).get("value"))
P
