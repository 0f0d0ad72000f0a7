
for name in main g0 s24 g0s24 s16 main; do
  if [ $name != main ]; then export LAKESOUL_B200_LIB=$PWD/build/variants/lib_$name.so; fi
  python bench.py --no-e2e --no-cpu --steps 10 --warmup 3 > gpurun_out/r2g_$name.json 2> gpurun_out/r2g_$name.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r2g_$name.json').read().strip().splitlines()[-1])
r=d['roofline']
print('$name', d['ms_per_step'], r['frac'], r['merge_ms'], r['gather_ms'])
"
done
