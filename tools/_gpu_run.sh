set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run6_pytest.log
tail -4 gpurun_out/r2_run6_pytest.log
for nc in 8 12 16 24; do
timeout 300 python bench.py --steps 8 --warmup 3 --contexts $nc --no-m2 --no-cpu-baseline > gpurun_out/r2_bench_v6_c${nc}.json 2>> gpurun_out/r2_bench_v6.err
done
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/_prove_once.py 3 > gpurun_out/r2_sanitizer_memcheck_prove.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer_memcheck_prove.log
tail -5 gpurun_out/r2_sanitizer_memcheck_prove.log
P2B_GRAPH=0 timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/_prove_once.py 1 > gpurun_out/r2_sanitizer_racecheck_prove.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2_sanitizer_racecheck_prove.log
tail -5 gpurun_out/r2_sanitizer_racecheck_prove.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_v6*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['e2e']['value'],1), round(d['e2e']['pageable_value'],1), d.get('launches_per_proof'), d['single_worker'].get('proofs_per_s'))
    except Exception as e: print(f, 'ERR', e)
PY
