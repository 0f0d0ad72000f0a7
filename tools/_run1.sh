LAKESOUL_B200_TILE_THREADS=256 python -m pytest tests/test_gpu_engine.py -m gpu -x -q 2>&1 | tail -2
for name in t512 t256 t512 t256; do
  if [ $name = t256 ]; then export LAKESOUL_B200_TILE_THREADS=256; else unset LAKESOUL_B200_TILE_THREADS; fi
  python bench.py --no-e2e --no-cpu --steps 10 --warmup 3 > gpurun_out/r2i_$name.json 2> gpurun_out/r2i_$name.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r2i_$name.json').read().strip().splitlines()[-1])
r=d['roofline']
print('$name', d['ms_per_step'], r['frac'], r['merge_ms'], r['gather_ms'])
"
done
