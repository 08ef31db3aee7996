set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_v9.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu_v9.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v9.json 2> gpurun_out/r2_bench_v9.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2_bench_v9.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_v9*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], round(d['value'],2), round(d['e2e']['value'],2), d['e2e'].get('pageable_value'), d.get('launches_per_proof'), d.get('single_worker'))
    except Exception as e: print(f, 'ERR', e)
PY
