mkdir -p gpurun_out
for v in v3 v5 v6 v6s3; do echo "== $v"; timeout 120 tools/build/pb_$v | grep -E "block=|checksum"; done > gpurun_out/pb_variants3.txt 2>&1
cat gpurun_out/pb_variants3.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for l in default v6mb3 v6s3 v3; do
  if [ $l = default ]; then unset P2B_LIB; else export P2B_LIB=$PWD/city_rollup_b200/build/libp2b_$l.so; fi
  timeout 200 python bench.py --steps 10 --warmup 3 --no-m2 --no-m1 --no-cpu-baseline > gpurun_out/bench_y_$l.json 2> gpurun_out/bench_y_$l.err
  echo $l; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_y_$l.json')); print(d['value'], d['ms_per_step'], d['stage_ms_per_step'], d['e2e']['value'])"
done
unset P2B_LIB
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
