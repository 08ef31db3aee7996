set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run5_pytest.log
tail -4 gpurun_out/r2_run5_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v5.json 2> gpurun_out/r2_bench_v5.err; echo "bench rc=$?"
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python bench.py --steps 6 --warmup 3 --contexts 16 --no-m2 --no-cpu-baseline > gpurun_out/r2_bench_v5_mc32_c16.json 2>> gpurun_out/r2_bench_v5.err
P2B_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,smsp__inst_executed.sum,launch__grid_size,launch__registers_per_thread --clock-control none --csv --log-file gpurun_out/r2_launches_prove_v5.csv python tools/_prove_once.py 3 > gpurun_out/r2_ncu_l_v5.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_v5*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['e2e']['value'],1), round(d['e2e']['pageable_value'],1), d.get('launches_per_proof'), d['single_worker'].get('proofs_per_s'), d['single_worker'].get('stage_ms_per_proof'))
    except Exception as e: print(f, 'ERR', e)
PY
