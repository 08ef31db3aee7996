set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v22.log 2>&1; tail -3 gpurun_out/pytest_gpu_v22.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_v22.json 2> gpurun_out/bench_v22.err; tail -2 gpurun_out/bench_v22.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_v22.json 2> gpurun_out/bench_ref_v22.err
