python -m pytest tests/test_gpu_full_size.py tests/test_gpu_reader.py -m gpu -q -x 2>&1 | tail -3
export SWEEP_PASSES=5
python tests/dev_slots_sweep.py 5 128 E3 24,32
python tests/dev_slots_sweep.py 2 32 E3 16
