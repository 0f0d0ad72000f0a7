for k in zstd_pages_kernel_v2:0 zstd_batch_kernel:0; do
  name=${k%%:*}; skip=${k##*:}
  ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 2 -o gpurun_out/r03_prof_$name -f \
      python tests/dev_one_bucket.py 5 1 2 E3 > gpurun_out/r03_prof_$name.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r03_launches_e2e_bucket.csv \
    python tests/dev_one_bucket.py 5 1 2 E3 > gpurun_out/r03_one_bucket.log 2>&1
tail -n 2 gpurun_out/r03_one_bucket.log
