python -m pytest tests/test_gpu_full_size.py tests/test_gpu_reader.py tests/test_gpu_file_groups.py -m gpu -q -x 2>&1 | tail -2
export SWEEP_PASSES=5
python tests/dev_slots_sweep.py 5 192 E3 16,24,32
python tests/dev_slots_sweep.py 2 32 E3 16
