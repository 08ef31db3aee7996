# ncu --set full of the quotient kernels of one City-shape proof (after a plain run of the same command)
set -x
python tools/_prove_once.py 1 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:k_quotient -c 8 -f -o gpurun_out/prof_quot_v22 python tools/_prove_once.py 1 > gpurun_out/ncu_full_quot_v22.log 2>&1
ncu -i gpurun_out/prof_quot_v22.ncu-rep --page raw --csv > gpurun_out/prof_quotient_v22_ncu_raw.csv 2>/dev/null
rm -f gpurun_out/prof_quot_v22.ncu-rep
wc -c gpurun_out/prof_quotient_v22_ncu_raw.csv
