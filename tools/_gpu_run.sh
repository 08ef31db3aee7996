set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v7.json 2> gpurun_out/r2_bench_v7.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2_bench_v7.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_v7_reference.json 2> gpurun_out/r2_bench_v7_reference.err; echo "ref rc=$?"
timeout 300 python tools/_quotient_bench.py 16 recursion 4 > gpurun_out/r2_quotient_2p16_recursion.txt 2>&1
P2B_QUOT_POINT_MAJOR=0 timeout 300 python tools/_quotient_bench.py 16 recursion 4 > gpurun_out/r2_quotient_2p16_recursion_gatemajor.txt 2>&1
timeout 300 python tools/_quotient_bench.py 16 city 4 > gpurun_out/r2_quotient_2p16_city.txt 2>&1
cat gpurun_out/r2_quotient_2p16_*.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,launch__registers_per_thread,sm__cycles_active.avg,sm__cycles_elapsed.max --clock-control none -k regex:"k_quotient" --csv --log-file gpurun_out/r2_quotient_2p16_ncu.csv python tools/_quotient_bench.py 16 recursion 1 > gpurun_out/r2_ncu_q.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__cycles_active.avg,sm__cycles_elapsed.max --clock-control none --csv --log-file gpurun_out/r2_commit_2p20x135_ncu.csv python tools/_commit_once.py 20 135 2 > gpurun_out/r2_ncu_c.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r2_commit_2p16x135_ncu.csv python tools/_commit_once.py 16 135 2 > gpurun_out/r2_ncu_c16.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_v7*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],2), round(d['e2e']['value'],2), d['e2e'].get('pageable_value'), d.get('launches_per_proof'), (d.get('single_worker') or {}).get('proofs_per_s'), (d.get('m2_lde_merkle_2p20x135') or {}).get('lde_merkle_ms'))
    except Exception as e: print(f, 'ERR', e)
PY
