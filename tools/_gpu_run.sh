for lib in libp2b.so libp2b_row2.so; do
echo "== $lib"
P2B_LIB=$PWD/city_rollup_b200/$lib timeout 300 python tools/_commit_once.py 16 135 3 2>&1 | tail -1
P2B_LIB=$PWD/city_rollup_b200/$lib timeout 300 python tools/_commit_once.py 20 135 3 2>&1 | tail -1
P2B_LIB=$PWD/city_rollup_b200/$lib timeout 300 python tools/_commit_farm.py 12 135 24 40 2>&1 | grep workers
P2B_LIB=$PWD/city_rollup_b200/$lib timeout 600 python bench.py --steps 6 --warmup 3 --no-m2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(d['value'],1), round(d['e2e']['value'],1))"
done
