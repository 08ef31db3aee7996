# usage: bash tools/qbench_run.sh G   (G = number of GPUs on the box) — builds the case and the tool, replays blocks
set -x
G=${1:-1}
python tools/dump_prove_case.py gpurun_out/prove_case.bin 12 2>&1 | tail -1
g++ -O2 -std=c++17 -I. tools/qbench_replay.cpp -Lcity_rollup_b200 -lp2b -lpthread -Wl,-rpath,$PWD/city_rollup_b200 -o tools/qbench_replay
for c in 1 8; do ./tools/qbench_replay -i gpurun_out/prove_case.bin -o gpurun_out/qbench_g${G}_c$c.json -n $((16 * G)) --gpus $G --contexts $c; done 2>&1 | tee gpurun_out/qbench_replay_${G}gpu.txt
rm -f gpurun_out/prove_case.bin
