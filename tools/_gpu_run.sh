set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_v13.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu_v13.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v13.json 2> gpurun_out/r2_bench_v13.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2_bench_v13.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_v13_reference.json 2> gpurun_out/r2_bench_v13_reference.err; echo "ref rc=$?"
timeout 300 python tools/_commit_once.py 16 135 4 > gpurun_out/r2_commit_2p16x135_v13.txt 2>&1
timeout 300 python tools/_commit_once.py 20 135 3 > gpurun_out/r2_commit_2p20x135_v13.txt 2>&1
timeout 300 python tools/_commit_once.py 20 400 3 > gpurun_out/r2_commit_2p20x400_v13.txt 2>&1
cat gpurun_out/r2_commit_2p*_v13.txt
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,launch__registers_per_thread,sm__cycles_active.avg,launch__grid_size --clock-control none --csv --log-file gpurun_out/r2_launches_prove_v13.csv python tools/_prove_once.py 3 > gpurun_out/r2_ncu_prove_v13.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_quotient_gates -c 4 -f -o gpurun_out/r2_prof_quotient_v13 python tools/_prove_once.py 1 > gpurun_out/r2_ncu_full_v13.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r2_prof_quotient_v13.ncu-rep --page raw --csv > gpurun_out/r2_prof_quotient_v13_ncu_raw.csv 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_v13*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], round(d['value'],2), round(d['e2e']['value'],2), d['e2e'].get('pageable_value'), d.get('launches_per_proof'), (d.get('single_worker') or {}).get('proofs_per_s'), (d.get('m2_lde_merkle_2p20x135') or {}).get('lde_merkle_ms'))
    except Exception as e: print(f, 'ERR', e)
PY
