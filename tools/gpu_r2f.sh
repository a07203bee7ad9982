#!/bin/bash
# round-2 GPU session F: decode_containers, new bench e2e, fast-math and tensor-core studies, launch lists for profiles/
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_demux_gpu.py tests/test_parity_lc_gpu.py -m gpu -q -x --timeout 600 > $O/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2f_pytest.log
tail -4 $O/r2f_pytest.log
( time python bench.py ) > $O/r2f_bench_full.log 2>&1; tail -c 300 $O/r2f_bench_full.log
python tools/fastmath_study.py --config 2 > $O/r2f_fast_c2_exact.json 2>$O/r2f_fast_err.log; cat $O/r2f_fast_c2_exact.json
JAADB200_LIB=$PWD/jaadec_b200/_build/variants/fmad.so python tools/fastmath_study.py --config 2 > $O/r2f_fast_c2_fmad.json 2>>$O/r2f_fast_err.log; cat $O/r2f_fast_c2_fmad.json
python tools/fastmath_study.py --config 3 --streams 32 > $O/r2f_fast_c3_exact.json 2>>$O/r2f_fast_err.log; cat $O/r2f_fast_c3_exact.json
JAADB200_LIB=$PWD/jaadec_b200/_build/variants/fmad.so python tools/fastmath_study.py --config 3 --streams 32 > $O/r2f_fast_c3_fmad.json 2>>$O/r2f_fast_err.log; cat $O/r2f_fast_c3_fmad.json
python tools/tc_qmf_study.py > $O/r2f_tc_qmf.json 2>$O/r2f_tc_err.log; cat $O/r2f_tc_qmf.json; tail -3 $O/r2f_tc_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r2f_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $O/r2f_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2f_launches_c3.csv python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2f_ncu3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2f_launches_c4.csv python bench.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r2f_ncu4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k2_filterbank|k1_parse|k2_prepass" -s 3 -c 3 -o $O/r2f_k1_k2_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $O/r2f_ncu2.log 2>&1
