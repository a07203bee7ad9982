#!/bin/bash
# session 3, call M: K3's process-error checks moved behind the warp shuffles (ptxas keeps the 72-register allocation):
# SBR / fuzz tests, then configs 3 and 4 A/B against the build from before this session (variant pre_pulse.so)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_fuzz_gpu.py tests/test_parity_sbr_gpu.py tests/test_parity_sbr_downsampled_gpu.py -q -m gpu -x > $O/r3m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r3m_pytest.log | cut -c1-300
for c in 3 4; do
  JAADB200_LIB=jaadec_b200/_build/variants/pre_pulse.so timeout 300 python bench.py --config $c --steps 3 --warmup 3 --no-e2e --no-extras --no-cpu-baseline > $O/r3m_c${c}_pre.json 2> $O/r3m_c${c}_pre.err; echo "pre $c rc=$?"
  timeout 300 python bench.py --config $c --steps 3 --warmup 3 --no-e2e --no-extras --no-cpu-baseline > $O/r3m_c${c}_now.json 2> $O/r3m_c${c}_now.err; echo "now $c rc=$?"
done
python - <<'P'
import json
for n in ("c3_pre","c3_now","c4_pre","c4_now"):
    d=json.load(open("gpurun_out/r3m_%s.json"%n)); print(n, round(d["value"]), d["ms_per_step"], d["roofline"]["kernel_ms"], d["bad_frames"], d.get("bad_frames_repeat_pass"))
P
