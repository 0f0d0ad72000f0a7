#!/bin/bash
# Profiles of one round (run on the GPU box through gpurun): launch list + DRAM bytes of one bench step, full ncu
# captures of the three step kernels that carry the time, launch list of one end-to-end bucket.
# usage: tools/collect_profiles.sh r03      -> gpurun_out/r03_*
R=${1:-rXX}
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${R}_launches_step_kernels.csv python bench.py --no-e2e --no-cpu --steps 2 --warmup 1 > gpurun_out/${R}_step.log 2>&1
for k in var_write_kernel:2 gather_fixed_kernel:2 tile_fast_kernel:8; do
  name=${k%%:*}; skip=${k##*:}
  ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -o gpurun_out/${R}_prof_$name -f \
      python bench.py --no-e2e --no-cpu --steps 2 --warmup 1 > gpurun_out/${R}_prof_$name.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${R}_launches_e2e_bucket.csv \
    python tests/dev_one_bucket.py 5 1 2 E3 > gpurun_out/${R}_one_bucket.log 2>&1
tail -n 2 gpurun_out/${R}_step.log; tail -n 2 gpurun_out/${R}_one_bucket.log
