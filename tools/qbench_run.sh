# usage: bash tools/qbench_run.sh G   (G = number of GPUs on the box) — builds the cases and the tool, replays
# (a) blocks shaped like qbench_data/example.bin, (b) BASELINE.json configs[4]: a binary aggregation tree over 2^10
# leaf proofs with 2^13-row circuits
set -x
G=${1:-1}
python tools/dump_prove_case.py gpurun_out/prove_case.bin 12 2>&1 | tail -1
python tools/dump_prove_case.py gpurun_out/prove_case13.bin 13 2>&1 | tail -1
g++ -O2 -std=c++17 -I. tools/qbench_replay.cpp -Lcity_rollup_b200 -lp2b -lpthread -Wl,-rpath,$PWD/city_rollup_b200 -o tools/qbench_replay
for c in 1 8; do ./tools/qbench_replay -i gpurun_out/prove_case.bin -o gpurun_out/qbench_g${G}_c$c.json -n $((16 * G)) --gpus $G --contexts $c; done 2>&1 | tee gpurun_out/qbench_replay_${G}gpu.txt
for c in 1 4; do ./tools/qbench_replay -i gpurun_out/prove_case13.bin -n 1 --agg-tree 10 --gpus $G --contexts $c; done 2>&1 | tee gpurun_out/agg_tree_${G}gpu.txt
rm -f gpurun_out/prove_case.bin gpurun_out/prove_case13.bin
