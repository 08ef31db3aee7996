set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_run1_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run1_pytest.log
tail -3 gpurun_out/r2_run1_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v1.json 2> gpurun_out/r2_bench_v1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_v1.err
for mc in 8 32; do for nc in 8 12 16; do
  CUDA_DEVICE_MAX_CONNECTIONS=$mc timeout 300 python bench.py --steps 6 --warmup 3 --contexts $nc --no-m2 --no-cpu-baseline > gpurun_out/r2_bench_v1_mc${mc}_c${nc}.json 2>> gpurun_out/r2_bench_v1_variants.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_v1*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['e2e']['value'],1), round(d['e2e']['pageable_value'],1), d.get('launches_per_proof'), d['single_worker'].get('proofs_per_s'))
    except Exception as e: print(f, 'ERR', e)
PY
