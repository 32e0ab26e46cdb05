set -x
# round-2 evidence run (one GPU): sanitizer on the new RANSAC paths, launch list of the bench, ncu captures of the two hot kernels
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ransac or pipeline or vote" > gpurun_out/r02_memcheck.log 2>&1; echo memcheck rc=$? >> gpurun_out/r02_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_filter_equals_fp32_scoring or tensor_filter_adversarial" > gpurun_out/r02_racecheck.log 2>&1; echo racecheck rc=$? >> gpurun_out/r02_racecheck.log
tail -4 gpurun_out/r02_memcheck.log gpurun_out/r02_racecheck.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-split-pair > gpurun_out/r02_launch_run.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ransac_kernel -s 2 -c 1 -o gpurun_out/r02_ransac_cfg2 python tools/ransac_bench.py 296 1 > gpurun_out/r02_ncu_cfg2.log 2>&1
BFR_CFG=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:ransac_kernel -s 2 -c 1 -o gpurun_out/r02_ransac_cfg3 python tools/ransac_bench.py 296 1 > gpurun_out/r02_ncu_cfg3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_tc_kernel -s 2 -c 1 -o gpurun_out/r02_k1tc python tools/k1_bench.py buffer_b200/libbuffer_b200.so 296 5000 1 planted > gpurun_out/r02_ncu_k1.log 2>&1
ls -la gpurun_out/r02_*
