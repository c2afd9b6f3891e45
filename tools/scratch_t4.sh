python -m pytest tests/test_gpu_dense_nms.py tests/test_gpu_fuzz.py tests/test_gpu_ultralytics.py tests/test_gpu_properties.py -x -q 2>&1 | tail -2
python tools/bench_configs.py --only 5 2>&1 | tail -1 | cut -c1-130
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__cycles_active.avg --clock-control none -k regex:k_dense_pairs -c 2 --csv python tools/bench_configs.py --only 5 --steps 3 2>/dev/null | grep -E "k_dense" | cut -d, -f5,15-
