# one context, one proof at a time, under latency-oriented settings of the tuning knobs
run() { echo "== $*"; env "$@" python bench.py --contexts 1 --steps 2 --warmup 3 --no-m2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['single_worker']; print(round(s['proofs_per_s'],1), round(s['ms_per_proof'],3), {k: round(v,3) for k,v in s['stage_ms_per_proof'].items()})"; }
run X=1
run P2B_POW_BLOCKS=148
run P2B_POW_BLOCKS=296
run P2B_POW_BLOCKS=148 P2B_TREE_FUSE_LOG=15
run P2B_POW_BLOCKS=148 P2B_TREE_FUSE_LOG=15 P2B_COOP_MAX_NODES=4096
run P2B_POW_BLOCKS=148 P2B_TREE_FUSE_LOG=15 P2B_HASH_BLOCK=128
