#!/bin/bash
# session 3, call P: default bench line and reference arm of the final build (bench.py after the results-only health check)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
python bench.py > $O/r3p_bench_default.json 2> $O/r3p_bench_default.err; echo "bench rc=$?"; cut -c1-200 $O/r3p_bench_default.json
python bench.py --impl reference > $O/r3p_bench_reference.json 2> $O/r3p_bench_reference.err; echo "ref rc=$?"; cut -c1-200 $O/r3p_bench_reference.json
