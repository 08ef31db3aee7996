# End-of-round GPU evidence: tests, plain bench, the reference arm, ncu launch lists of the same commands.
# (The ncu --set full captures of the hot kernels were taken when those kernels last changed: profiles/r01_prof_*.)
set -x
V=${1:-v11}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; tail -3 gpurun_out/pytest_gpu_final.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$V.json 2> gpurun_out/bench_$V.err; tail -2 gpurun_out/bench_$V.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$V.json 2> gpurun_out/bench_ref_$V.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$V.csv python bench.py --steps 2 --warmup 3 --no-m2 --no-m1 --no-cpu-baseline > gpurun_out/ncu_l_$V.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/prove_launches_$V.csv python tools/prove_once.py > gpurun_out/ncu_prove_$V.log 2>&1
python tools/prove_trace.py 2>&1 | tail -10 > gpurun_out/prove_trace_$V.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
# one ncu --set full capture of the dominant kernel (after the plain runs above), raw page exported next to it
ncu --set full --clock-control none --import-source on -k regex:k_leaf_hash_colmajor -c 1 -f -o gpurun_out/prof_leaf_$V python bench.py --steps 1 --warmup 1 --no-m2 --no-m1 --no-cpu-baseline > gpurun_out/ncu_full_$V.log 2>&1
ncu -i gpurun_out/prof_leaf_$V.ncu-rep --page raw --csv > gpurun_out/prof_leaf_${V}_ncu_raw.csv 2>/dev/null
