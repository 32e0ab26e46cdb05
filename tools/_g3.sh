set -x
BFR_SO=variants/lib_rt.so python tools/ransac_bench.py 1623 1 2 4 8 > gpurun_out/r2c_rs_cfg2.log 2>&1
BFR_SO=variants/lib_rt.so BFR_CFG=3 python tools/ransac_bench.py 1623 1 2 4 > gpurun_out/r2c_rs_cfg3.log 2>&1
python tools/ransac_bench.py 203 1 2 3 4 6 8 12 16 > gpurun_out/r2c_rs_cfg2_203.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ransac_kernel -s 2 -c 1 -o gpurun_out/r2c_ransac_cfg2 python tools/ransac_bench.py 296 1 > gpurun_out/r2c_ncu_cfg2.log 2>&1
BFR_CFG=3 ncu --set full --clock-control none --import-source on -k regex:ransac_kernel -s 2 -c 1 -o gpurun_out/r2c_ransac_cfg3 python tools/ransac_bench.py 296 1 > gpurun_out/r2c_ncu_cfg3.log 2>&1
grep -h "splits\|per CTA" gpurun_out/r2c_rs_*.log
