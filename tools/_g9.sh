set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k1_|ransac|post_refinement|rigid|select|prep|compact|decode" -s 24 -c 18 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-split-pair > gpurun_out/r02_launch_run.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ransac_kernel -s 2 -c 1 -o gpurun_out/r02_ransac_cfg2 -f python tools/ransac_bench.py 1623 1 > gpurun_out/r02_ncu_cfg2.log 2>&1
BFR_CFG=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:ransac_kernel -s 2 -c 1 -o gpurun_out/r02_ransac_cfg3 -f python tools/ransac_bench.py 1623 1 > gpurun_out/r02_ncu_cfg3.log 2>&1
BFR_SO=variants/lib_tr.so timeout 100 python tools/ransac_trace.py > gpurun_out/r02_trace.log 2>&1
BFR_SO=variants/lib_rt.so timeout 150 python tools/ransac_bench.py 1623 1 > gpurun_out/r02_rt_cfg2.log 2>&1
BFR_CFG=3 BFR_SO=variants/lib_rt.so timeout 150 python tools/ransac_bench.py 1623 1 > gpurun_out/r02_rt_cfg3.log 2>&1
python tools/latency_one_pair.py > gpurun_out/r02_latency.log 2>&1
tail -3 gpurun_out/r02_trace.log gpurun_out/r02_rt_cfg2.log gpurun_out/r02_rt_cfg3.log gpurun_out/r02_latency.log
