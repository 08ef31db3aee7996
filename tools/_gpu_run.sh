set -x
timeout 300 tools/poseidon_bench > gpurun_out/r2_poseidon_bench_q.txt 2>&1; tail -30 gpurun_out/r2_poseidon_bench_q.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_leaf_hash_colmajor -c 1 -f -o gpurun_out/r2_prof_leaf_q python tools/_commit_once.py 16 135 1 > gpurun_out/r2_ncu_full_leaf_q.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r2_prof_leaf_q.ncu-rep --page raw --csv > gpurun_out/r2_prof_leaf_q_ncu_raw.csv 2>/dev/null
rm -f gpurun_out/r2_prof_leaf_q.ncu-rep
