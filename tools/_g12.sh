set -x
timeout 120 python tools/ransac_tc_check.py 296 > gpurun_out/r3k_tc_check.log 2>&1; tail -2 gpurun_out/r3k_tc_check.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_filter_equals_fp32_scoring or ransac_counts" > gpurun_out/r02_synccheck.log 2>&1; echo synccheck rc=$? >> gpurun_out/r02_synccheck.log; grep "=========" gpurun_out/r02_synccheck.log | grep -v "Host Frame" | head -8; tail -3 gpurun_out/r02_synccheck.log
