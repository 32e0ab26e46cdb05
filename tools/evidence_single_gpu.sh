set -x
timeout 120 python tools/ransac_tc_check.py 296 > gpurun_out/r3o_tc_check.log 2>&1; tail -2 gpurun_out/r3o_tc_check.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r3c_tests.log 2>&1; tail -2 gpurun_out/r3c_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k1_|ransac|post_refinement|rigid|select|prep|compact|decode" -s 24 -c 18 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-split-pair > gpurun_out/r02_launch_run.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ransac_kernel -s 2 -c 1 -o gpurun_out/r02_ransac_cfg2 -f python tools/ransac_bench.py 1623 1 > gpurun_out/r02_ncu_cfg2.log 2>&1
BFR_CFG=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:ransac_kernel -s 2 -c 1 -o gpurun_out/r02_ransac_cfg3 -f python tools/ransac_bench.py 1623 1 > gpurun_out/r02_ncu_cfg3.log 2>&1
BFR_SO=variants/lib_tr.so timeout 100 python tools/ransac_trace.py > gpurun_out/r02_trace.log 2>&1
BFR_SO=variants/lib_rt.so timeout 150 python tools/ransac_bench.py 1623 1 > gpurun_out/r02_rt_cfg2.log 2>&1
BFR_CFG=3 BFR_SO=variants/lib_rt.so timeout 150 python tools/ransac_bench.py 1623 1 > gpurun_out/r02_rt_cfg3.log 2>&1
python tools/latency_one_pair.py > gpurun_out/r02_latency.log 2>&1
timeout 400 python bench.py > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ransac or pipeline or vote or full_size_batch" > gpurun_out/r02_memcheck.log 2>&1; echo memcheck rc=$? >> gpurun_out/r02_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_filter_equals_fp32_scoring or tensor_filter_adversarial" > gpurun_out/r02_racecheck.log 2>&1; echo racecheck rc=$? >> gpurun_out/r02_racecheck.log
cat gpurun_out/r02_rt_cfg2.log gpurun_out/r02_rt_cfg3.log gpurun_out/r02_latency.log; tail -3 gpurun_out/r02_memcheck.log gpurun_out/r02_racecheck.log
