set -x
python -m pytest tests/test_gpu_prove.py -m gpu -x -q -k "nowait or one_thread or refilled" 2>&1 | tail -3
python tools/dump_prove_case.py gpurun_out/prove_case.bin 12 2>&1 | tail -1
g++ -O2 -std=c++17 -I. tools/qbench_replay.cpp -Lcity_rollup_b200 -lp2b -lpthread -Wl,-rpath,$PWD/city_rollup_b200 -o tools/qbench_replay
for k in 8 12 16; do ./tools/qbench_replay -i gpurun_out/prove_case.bin -d tests/golden/example_dag.bin -n 16 --gpus 1 --async $k; done 2>&1 | tee gpurun_out/qbench_async_v21.txt | cut -c1-120,330-520
rm -f gpurun_out/prove_case.bin
