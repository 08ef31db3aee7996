# usage: bash tools/qbench_run.sh G   (G = number of GPUs on the box) — builds the cases and the tool, replays
# (a) the job DAG of qbench_data/example.bin (read from tests/golden/example_dag.bin: counters, goals, next-job lists;
#     43 plonky2 jobs / 67 proofs per block), worker threads with blocking p2b_prove and one host thread per GPU driving
#     its contexts through p2b_prove_submit / collect,
# (b) BASELINE.json configs[4]: a binary aggregation tree over 2^10 leaf proofs with 2^13-row circuits
set -x
G=${1:-1}
python tools/dump_prove_case.py gpurun_out/prove_case.bin 12 2>&1 | tail -1
python tools/dump_prove_case.py gpurun_out/prove_case13.bin 13 2>&1 | tail -1
g++ -O2 -std=c++17 -I. tools/qbench_replay.cpp -Lcity_rollup_b200 -lp2b -lpthread -Wl,-rpath,$PWD/city_rollup_b200 -o tools/qbench_replay
for c in 1 8 24; do ./tools/qbench_replay -i gpurun_out/prove_case.bin -d tests/golden/example_dag.bin -o gpurun_out/qbench_g${G}_c$c.json -n $((16 * G)) --gpus $G --contexts $c; done 2>&1 | tee gpurun_out/qbench_replay_${G}gpu.txt
./tools/qbench_replay -i gpurun_out/prove_case.bin -d tests/golden/example_dag.bin -n $((16 * G)) --gpus $G --async 24 2>&1 | tee -a gpurun_out/qbench_replay_${G}gpu.txt
for c in 1 4 16; do ./tools/qbench_replay -i gpurun_out/prove_case13.bin -n 1 --agg-tree 10 --gpus $G --contexts $c; done 2>&1 | tee gpurun_out/agg_tree_${G}gpu.txt
rm -f gpurun_out/prove_case.bin gpurun_out/prove_case13.bin
