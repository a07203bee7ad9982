#!/bin/bash
# round-2 GPU session J: source-level profile of K3 (configs 3 and 4)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for c in 4 3; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k3_sbr" -c 1 -o $O/r2j_k3_c$c -f python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2j_ncu_c$c.log 2>&1
  echo "ncu c$c rc=$?"
done
ls -la $O | grep r2j
