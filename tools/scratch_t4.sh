python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_tick.py tests/test_gpu_ultralytics.py tests/test_gpu_properties.py tests/test_collector.py -x -q > gpurun_out/post_test.log 2>&1; tail -4 gpurun_out/post_test.log
python tools/phase_timing_fused.py 2>&1 | tail -5
python tools/bench_configs.py --only 1,2,5 > gpurun_out/cfg12.log 2>&1; tail -3 gpurun_out/cfg12.log | cut -c1-330
