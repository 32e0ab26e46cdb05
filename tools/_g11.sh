set -x
timeout 120 python tools/ransac_tc_check.py 296 > gpurun_out/r3j_tc_check.log 2>&1; tail -2 gpurun_out/r3j_tc_check.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_filter_equals_fp32_scoring or ransac_counts" > gpurun_out/r02_synccheck.log 2>&1; echo synccheck rc=$? >> gpurun_out/r02_synccheck.log; grep "=========" gpurun_out/r02_synccheck.log | head -5; tail -3 gpurun_out/r02_synccheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mutual_nn_bit_exact" > gpurun_out/r02_synccheck_k1.log 2>&1; echo synccheck rc=$? >> gpurun_out/r02_synccheck_k1.log; grep "=========" gpurun_out/r02_synccheck_k1.log | head -5; tail -3 gpurun_out/r02_synccheck_k1.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r3j_tests.log 2>&1; tail -2 gpurun_out/r3j_tests.log
python __graft_entry__.py smoke 2>&1 | tail -2
for c in 2 3; do BFR_CFG=$c timeout 150 python tools/ransac_bench.py 1623 1 2>&1 | head -1; done
