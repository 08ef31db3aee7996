timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-m2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(d['value'],1), round(d['e2e']['value'],1), d['single_worker']['proofs_per_s'])"
timeout 300 python tools/_quotient_bench.py 16 city 3 2>&1 | tail -1
timeout 300 python tools/_quotient_bench.py 16 recursion 3 2>&1 | tail -1
