python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_ultralytics.py -x -q -k "letterbox or preprocess or resize or geom or fused or motion" > gpurun_out/lb_test.log 2>&1; tail -3 gpurun_out/lb_test.log
python tools/bench_configs.py --only L 2>&1 | tail -12 | cut -c1-200
