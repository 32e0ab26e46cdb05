set -x
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
tail -c 2500 gpurun_out/r2d_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d_bench.json'))
for k in ('value','ms_per_step','scaling','e2e','split_pair','cpu_baseline','cpu_baseline_torch'): print(k, d.get(k))
print(d['roofline_ransac']['open3d_confidence_exit'])
PY
