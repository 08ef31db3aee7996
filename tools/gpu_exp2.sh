mkdir -p gpurun_out
for v in "$@"; do echo "== $v"; timeout 120 tools/build/pb_$v | grep -E "block=|checksum|^stream"; done > gpurun_out/pb_variants4.txt 2>&1
cat gpurun_out/pb_variants4.txt
