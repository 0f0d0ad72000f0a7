python -m pytest tests/test_gpu_engine.py tests/test_gpu_reader.py -m gpu -x -q 2>&1 | tail -3
for v in "main:X=1" "pf2:X=1"; do
  IFS=: read name e1 <<< "$v"
  if [ $name = pf2 ]; then export LAKESOUL_B200_LIB=$PWD/build/variants/lib_pf2.so; fi
  env $e1 python bench.py --no-e2e --no-cpu --steps 10 --warmup 3 > gpurun_out/r2f_$name.json 2> gpurun_out/r2f_$name.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r2f_$name.json').read().strip().splitlines()[-1])
r=d['roofline']
print('$name', d['ms_per_step'], r['frac'], r['merge_ms'], r['gather_ms'])
"
done
