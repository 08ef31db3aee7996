# End-of-round GPU evidence: tests, plain bench, ncu launch list of the same command, one full capture of the new kernel.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; tail -3 gpurun_out/pytest_gpu_final.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v10.json 2> gpurun_out/bench_v10.err; tail -2 gpurun_out/bench_v10.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_v10.json 2> gpurun_out/bench_ref_v10.err; cat gpurun_out/bench_ref_v10.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v10.csv python bench.py --steps 2 --warmup 3 --no-m2 --no-m1 --no-cpu-baseline > gpurun_out/ncu_l10.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_leaf_absorb_colmajor --launch-skip 6 --launch-count 1 -o gpurun_out/prof_absorb -f python bench.py --steps 2 --warmup 3 --no-m2 --no-m1 --no-cpu-baseline > gpurun_out/ncu_f10.log 2>&1
ncu -i gpurun_out/prof_absorb.ncu-rep --page raw --csv > gpurun_out/prof_absorb_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/prove_launches_v3.csv python tools/prove_once.py > gpurun_out/ncu_prove3.log 2>&1
ls -la gpurun_out | tail -12
