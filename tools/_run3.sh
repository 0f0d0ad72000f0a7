for f in build/variants/lib_*.so; do
  name=$(basename $f .so)
  LAKESOUL_B200_LIB=$PWD/$f python bench.py --no-e2e --no-cpu --steps 10 --warmup 3 > gpurun_out/r2e_$name.json 2> gpurun_out/r2e_$name.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r2e_$name.json').read().strip().splitlines()[-1])
r=d['roofline']
print('$name', round(d['ms_per_step'],3), round(r['frac'],4), round(r['merge_ms'],3), round(r['gather_ms'],3))
"
done
