python -m pytest tests/test_gpu_writer.py -m gpu -x -q 2>&1 | tail -25

