set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v5.log 2>&1; tail -3 gpurun_out/pytest_gpu_v5.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; tail -2 gpurun_out/bench_v5.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v5.csv python bench.py --steps 2 --warmup 3 --no-m2 --no-m1 --no-cpu-baseline > gpurun_out/ncu_l5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_leaf_hash_colmajor --launch-skip 3 --launch-count 1 -o gpurun_out/prof_leaf_v3 -f python bench.py --steps 2 --warmup 3 --no-m2 --no-m1 --no-cpu-baseline > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_row4096 --launch-skip 6 --launch-count 1 -o gpurun_out/prof_row4096 -f python bench.py --steps 2 --warmup 3 --no-m2 --no-m1 --no-cpu-baseline > gpurun_out/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_strided --launch-skip 7 --launch-count 1 -o gpurun_out/prof_strided -f python bench.py --steps 2 --warmup 3 --no-m2 --no-m1 --no-cpu-baseline > gpurun_out/ncu_f3.log 2>&1
for f in prof_leaf_v3 prof_row4096 prof_strided; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null; done
ls -la gpurun_out | tail -12
