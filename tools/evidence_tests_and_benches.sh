set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r3c_tests.log 2>&1; tail -3 gpurun_out/r3c_tests.log
timeout 400 python bench.py > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err; tail -c 300 gpurun_out/r3c_bench.err
timeout 300 python bench.py --config 3 --steps 2 --warmup 3 --no-e2e --no-cpu --no-split-pair > gpurun_out/r3c_bench_cfg3.json 2> gpurun_out/r3c_bench_cfg3.err
timeout 300 python bench.py --config 4 --steps 2 --warmup 3 --no-e2e --no-cpu --no-split-pair > gpurun_out/r3c_bench_cfg4.json 2> gpurun_out/r3c_bench_cfg4.err
python - <<PY
import json
for f in ("r3c_bench","r3c_bench_cfg3","r3c_bench_cfg4"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
        print(f, round(d["value"]), round(d["ms_per_step"],2), round(d["roofline"]["k1_ms_per_launch"],2), round(d["roofline_ransac"]["ms_per_launch"],2), d["roofline_ransac"].get("ms_per_launch_fp32_scoring"), d["roofline_ransac"].get("outputs_identical_to_fp32_scoring"), d["quality"]["registration_recall"], d.get("e2e",{}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
