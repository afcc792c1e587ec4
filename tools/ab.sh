python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "w8" 2>&1 | tail -12
