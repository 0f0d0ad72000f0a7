
for name in main main; do
  python bench.py --no-e2e --no-cpu --steps 10 --warmup 3 > gpurun_out/r2h_$name.json 2> gpurun_out/r2h_$name.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r2h_$name.json').read().strip().splitlines()[-1])
r=d['roofline']
print('$name', d['ms_per_step'], r['frac'], r['merge_ms'], r['gather_ms'])
"
done
