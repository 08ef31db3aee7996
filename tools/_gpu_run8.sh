# one process, G GPUs: the qbench-shaped replay (dumped DAG) and the aggregation tree; usage: bash tools/_gpu_run8.sh G
set -x
G=${1:-8}
python tools/dump_prove_case.py gpurun_out/prove_case.bin 12 2>&1 | tail -1
python tools/dump_prove_case.py gpurun_out/prove_case13.bin 13 2>&1 | tail -1
g++ -O2 -std=c++17 -I. tools/qbench_replay.cpp -Lcity_rollup_b200 -lp2b -lpthread -Wl,-rpath,$PWD/city_rollup_b200 -o tools/qbench_replay
nproc
./tools/qbench_replay -i gpurun_out/prove_case.bin -d tests/golden/example_dag.bin -n $((24 * G)) --gpus $G --contexts 24 2>&1 | tee gpurun_out/qbench_replay_${G}gpu_v21.txt
./tools/qbench_replay -i gpurun_out/prove_case.bin -d tests/golden/example_dag.bin -n $((24 * G)) --gpus $G --contexts 16 2>&1 | tee -a gpurun_out/qbench_replay_${G}gpu_v21.txt
./tools/qbench_replay -i gpurun_out/prove_case.bin -d tests/golden/example_dag.bin -n $((24 * G)) --gpus $G --async 8 2>&1 | tee -a gpurun_out/qbench_replay_${G}gpu_v21.txt
./tools/qbench_replay -i gpurun_out/prove_case13.bin -n 1 --agg-tree 10 --gpus $G --contexts 8 2>&1 | tee gpurun_out/agg_tree_${G}gpu_v21.txt
./tools/qbench_replay -i gpurun_out/prove_case13.bin -n 1 --agg-tree 10 --gpus $G --contexts 16 2>&1 | tee -a gpurun_out/agg_tree_${G}gpu_v21.txt
rm -f gpurun_out/prove_case.bin gpurun_out/prove_case13.bin
