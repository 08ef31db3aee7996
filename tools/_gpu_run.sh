python tools/dump_prove_case.py gpurun_out/prove_case.bin 12 2>&1 | tail -1
python tools/dump_prove_case.py gpurun_out/prove_case13.bin 13 2>&1 | tail -1
g++ -O2 -std=c++17 -I. tools/qbench_replay.cpp -Lcity_rollup_b200 -lp2b -lpthread -Wl,-rpath,$PWD/city_rollup_b200 -o tools/qbench_replay
( for c in 1 8 24; do ./tools/qbench_replay -i gpurun_out/prove_case.bin -d tests/golden/example_dag.bin -o gpurun_out/qbench_g1_c$c.json -n 64 --contexts $c 2>&1 | tail -1; done
for k in 8 16; do ./tools/qbench_replay -i gpurun_out/prove_case.bin -d tests/golden/example_dag.bin -n 64 --async $k 2>&1 | tail -1; done
for c in 1 8 24; do ./tools/qbench_replay -i gpurun_out/prove_case13.bin -n 1 --agg-tree 10 --contexts $c 2>&1 | tail -1; done ) > gpurun_out/r2_qbench_replay_1gpu.txt
cat gpurun_out/r2_qbench_replay_1gpu.txt | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(d['harness'][:30], d['mode'][:32], 'ctx', d['contexts_per_gpu'], 'blocks', d['blocks'], 'proofs', d['proofs'], 'wall', d['wall_s'], 'proofs/s', d['proofs_per_s'], 'busy', d.get('worker_busy_fraction'), 'mismatch', d['mismatching_proofs'])
"
head -c 600 gpurun_out/qbench_g1_c24.json
rm -f gpurun_out/prove_case.bin gpurun_out/prove_case13.bin
