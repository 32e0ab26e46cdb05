set -x
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg2_${N}gpu.json 2> gpurun_out/r02_bench_cfg2_${N}gpu.err
tail -c 1500 gpurun_out/r02_bench_cfg2_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_cfg2_${N}gpu.json') if l.startswith('{')][-1])
for k in ('value','ms_per_step','scaling','weak_scaling','e2e','split_pair'): print(k, d.get(k))
print(d['roofline']['k1_ms_per_launch'], d['roofline_ransac']['ms_per_launch'])
PY
