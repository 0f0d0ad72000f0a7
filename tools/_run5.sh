export SWEEP_PASSES=5
echo "== lanes"; python tests/dev_slots_sweep.py 5 256 E3 24
echo "== warps"; LAKESOUL_B200_ZSTD_PLAN_WARPS=1 python tests/dev_slots_sweep.py 5 256 E3 24
echo "== lanes"; python tests/dev_slots_sweep.py 5 256 E3 24,32
echo "== warps"; LAKESOUL_B200_ZSTD_PLAN_WARPS=1 python tests/dev_slots_sweep.py 5 256 E3 24
nproc
