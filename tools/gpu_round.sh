# End-of-round GPU evidence (one B200): tests, plain bench, the reference arm, the ncu launch list of three proofs, one
# ncu --set full capture of the dominant kernel, the commit stage timers.  usage: bash tools/gpu_round.sh v18
# (bench.py itself does not complete under ncu: 24 host threads and graph replays under kernel serialisation.)
set -x
V=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$V.log 2>&1; tail -3 gpurun_out/pytest_gpu_$V.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_$V.json 2> gpurun_out/bench_$V.err; tail -2 gpurun_out/bench_$V.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$V.json 2> gpurun_out/bench_ref_$V.err
for s in "16 135 3" "20 135 3" "20 400 2"; do python tools/_commit_once.py $s; done > gpurun_out/commit_stages_$V.txt 2>&1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,launch__registers_per_thread,sm__cycles_active.avg,launch__grid_size --clock-control none --csv --log-file gpurun_out/launches_prove_$V.csv python tools/_prove_once.py 3 > gpurun_out/ncu_prove_$V.log 2>&1
python tools/launch_summary.py gpurun_out/launches_prove_$V.csv 3 | head -12
# one ncu --set full capture of the dominant kernel (after the plain runs above), raw page exported next to it
ncu --set full --clock-control none --import-source on -k regex:k_leaf_hash_colmajor -c 1 -f -o gpurun_out/prof_leaf_$V python tools/_commit_once.py 16 135 1 > gpurun_out/ncu_full_$V.log 2>&1
ncu -i gpurun_out/prof_leaf_$V.ncu-rep --page raw --csv > gpurun_out/prof_leaf_${V}_ncu_raw.csv 2>/dev/null
rm -f gpurun_out/prof_leaf_$V.ncu-rep
