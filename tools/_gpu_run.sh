set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_v15.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu_v15.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_v15.json 2> gpurun_out/r2_bench_v15.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2_bench_v15.err
timeout 300 python tools/_commit_once.py 16 135 4 > gpurun_out/r2_commit_2p16x135_v15.txt 2>&1
cat gpurun_out/r2_commit_2p16x135_v15.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_v15*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], round(d['value'],2), round(d['e2e']['value'],2), d['e2e'].get('pageable_value'), d.get('launches_per_proof'), (d.get('single_worker') or {}).get('proofs_per_s'), (d.get('m2_lde_merkle_2p20x135') or {}))
    except Exception as e: print(f, 'ERR', e)
PY
