set -x
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ransac or pipeline or vote or full_size_batch" > gpurun_out/r02_memcheck.log 2>&1; echo memcheck rc=$? >> gpurun_out/r02_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_filter_equals_fp32_scoring or tensor_filter_adversarial" > gpurun_out/r02_racecheck.log 2>&1; echo racecheck rc=$? >> gpurun_out/r02_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_filter_equals_fp32_scoring or ransac_counts" > gpurun_out/r02_synccheck.log 2>&1; echo synccheck rc=$? >> gpurun_out/r02_synccheck.log
cat gpurun_out/r02_memcheck.log gpurun_out/r02_racecheck.log gpurun_out/r02_synccheck.log | tail -30
