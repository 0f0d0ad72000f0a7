LAKESOUL_B200_ZSTD_PROFILE=1 LAKESOUL_B200_ZSTD_STREAMS=1 python tests/dev_one_bucket.py 5 1 3 E3 2>&1 | grep -v "^rep 0" | tail -12
