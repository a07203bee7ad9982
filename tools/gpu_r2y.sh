#!/bin/bash
# round-2 GPU session Y: the parity suite's SBR / PS / LC / corrupted-stream cases against a build with device-side bounds
# asserts (-DJAADB_BOUNDS_ASSERT, tools/build_variants.sh) on the new shared-memory windows and prefetch ranges.
# (compute-sanitizer is closed on this pool; this is the check its refusal message asks for instead.)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
JAADB200_LIB=jaadec_b200/_build/variants/asserts.so timeout 1200 python -m pytest tests/test_parity_sbr_gpu.py tests/test_parity_sbr_downsampled_gpu.py tests/test_parity_lc_gpu.py tests/test_fuzz_gpu.py -m gpu -q -x --timeout 900 > $O/r2y_asserts_pytest.log 2>&1; echo "asserts build pytest rc=$?" >> $O/r2y_asserts_pytest.log
tail -4 $O/r2y_asserts_pytest.log
grep -c -i "assert" $O/r2y_asserts_pytest.log
