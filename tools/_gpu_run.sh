set -x
mkdir -p gpurun_out
timeout 300 python - > gpurun_out/r2_run3_plan.txt 2>&1 <<'PY'
import sys, time
sys.path[:0]=['.','tests','tools']
import numpy as np
import city_rollup_b200 as m, prove_bench as PB
circ,digest,pis=PB.build_case()
c=m.Context(0)
cd=m.CircuitData(c,circ.desc()); cs=m.PolynomialBatch.from_values(c,circ.constants_sigmas_values(),3,False,4,keep_values=True)
params=m.FriParams(3,4,16,28,[4,4]); wv=np.stack(circ.wire_values())
outs=[]
for i in range(6):
    l0=c.launch_count(); t0=time.perf_counter()
    outs.append(m.prove_native(c,cd,cs,digest,wv,pis,params,raw=True))
    print("proof", i, "launches", c.launch_count()-l0, "ms", round((time.perf_counter()-t0)*1e3,3), "plan", c.plan_info())
print("all equal", all((o==outs[0]).all() for o in outs))
t0=time.perf_counter()
for i in range(50): m.prove_native(c,cd,cs,digest,wv,pis,params,raw=True)
print("ms per proof (graph, 1 ctx)", (time.perf_counter()-t0)/50*1e3)
PY
cat gpurun_out/r2_run3_plan.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run3_pytest.log
tail -5 gpurun_out/r2_run3_pytest.log
P2B_GRAPH=0 timeout 900 python -m pytest tests/test_gpu_prove.py -m gpu -x -q > gpurun_out/r2_run3_pytest_nograph.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run3_pytest_nograph.log
tail -3 gpurun_out/r2_run3_pytest_nograph.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v3.json 2> gpurun_out/r2_bench_v3.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2_bench_v3.err
P2B_GRAPH=0 timeout 300 python bench.py --steps 6 --warmup 3 --no-m2 --no-cpu-baseline > gpurun_out/r2_bench_v3_nograph.json 2>> gpurun_out/r2_bench_v3.err
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python bench.py --steps 6 --warmup 3 --contexts 16 --no-m2 --no-cpu-baseline > gpurun_out/r2_bench_v3_mc32_c16.json 2>> gpurun_out/r2_bench_v3.err
timeout 300 python bench.py --steps 6 --warmup 3 --contexts 12 --no-m2 --no-cpu-baseline > gpurun_out/r2_bench_v3_c12.json 2>> gpurun_out/r2_bench_v3.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_v3*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['e2e']['value'],1), round(d['e2e']['pageable_value'],1), d.get('launches_per_proof'), d['single_worker'].get('proofs_per_s'), d['e2e']['host_cpu_ms_per_proof'])
    except Exception as e: print(f, 'ERR', e)
PY
