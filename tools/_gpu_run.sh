set -x
python -m pytest tests/test_gpu_prove.py -m gpu -x -q -k "latency or tuning or nowait" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-m2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['single_worker']; print(round(d['value'],1), round(d['e2e']['value'],1), round(s['proofs_per_s'],1), s['latency_mode'])"
P2B_MODE=latency python bench.py --steps 5 --warmup 3 --no-m2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('all workers in latency mode:', round(d['value'],1))"
