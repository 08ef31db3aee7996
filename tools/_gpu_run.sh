set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_v12.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_gpu_v12.log
timeout 300 python tools/_quotient_bench.py 12 city 4 > gpurun_out/r2_quotient_2p12_city_v12.txt 2>&1
timeout 300 python tools/_quotient_bench.py 16 recursion 4 > gpurun_out/r2_quotient_2p16_recursion_v12.txt 2>&1
tail -n 2 gpurun_out/r2_quotient_2p1*_v12.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-m2 > gpurun_out/r2_bench_v12.json 2> gpurun_out/r2_bench_v12.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2_bench_v12.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_v12*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], round(d['value'],2), round(d['e2e']['value'],2), d['e2e'].get('pageable_value'), d.get('launches_per_proof'), d.get('single_worker'))
    except Exception as e: print(f, 'ERR', e)
PY
