python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_parity.py -x -q -k "grid_paths or dense_long_lived" 2>&1 | tail -3
