python -m pytest tests/test_gpu_full_size.py tests/test_gpu_reader.py -m gpu -q -x 2>&1 | tail -2
LAKESOUL_B200_ZSTD_PROFILE=1 LAKESOUL_B200_ZSTD_STREAMS=1 python tests/dev_one_bucket.py 5 1 2 E3 2>&1 | grep -v "^rep 0" | grep "ZSTDPROF2b\|batch"
