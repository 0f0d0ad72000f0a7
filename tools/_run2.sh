ncu --set full --clock-control none --import-source on -k regex:var_write_kernel -s 2 -c 1 -o gpurun_out/r2c_prof_varw -f python bench.py --no-e2e --no-cpu --steps 2 --warmup 1 > gpurun_out/r2c_prof.log 2>&1
tail -3 gpurun_out/r2c_prof.log
