set -x
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_v18_${N}gpu.json 2> gpurun_out/r2_bench_v18_${N}gpu.err; echo "bench$N rc=$?"
tail -c 400 gpurun_out/r2_bench_v18_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r2_bench_v18_${N}gpu_reference.json 2> gpurun_out/r2_bench_v18_${N}gpu_reference.err; echo "ref$N rc=$?"
nproc
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_v18_${N}gpu*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], round(d['value'],2), round(d['e2e']['value'],2), d['e2e'].get('pageable_value'), d.get('launches_per_proof'), d.get('cpu_baseline',{}).get('cores'))
    except Exception as e: print(f, 'ERR', e)
PY
