set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_tests.log 2>&1
tail -n 30 gpurun_out/r2b_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -c 3000 gpurun_out/r2b_bench.json
timeout 300 python bench.py --config 3 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2b_bench_cfg3.json 2> gpurun_out/r2b_bench_cfg3.err
tail -c 1500 gpurun_out/r2b_bench_cfg3.json
