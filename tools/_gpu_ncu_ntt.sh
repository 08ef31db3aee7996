# ncu --set full of the NTT / LDE kernels of one 2^16 x 135 commit (after a plain run of the same command)
set -x
python tools/_commit_once.py 16 135 2 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:k_row4096 -c 2 -f -o gpurun_out/prof_ntt_v22 python tools/_commit_once.py 16 135 1 > gpurun_out/ncu_full_ntt_v22.log 2>&1
ncu -i gpurun_out/prof_ntt_v22.ncu-rep --page raw --csv > gpurun_out/prof_ntt_row_v22_ncu_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_strided -c 2 -f -o gpurun_out/prof_ntt_v22 python tools/_commit_once.py 16 135 1 >> gpurun_out/ncu_full_ntt_v22.log 2>&1
ncu -i gpurun_out/prof_ntt_v22.ncu-rep --page raw --csv > gpurun_out/prof_ntt_strided_v22_ncu_raw.csv 2>/dev/null
rm -f gpurun_out/prof_ntt_v22.ncu-rep
wc -c gpurun_out/prof_ntt_*_v22_ncu_raw.csv
