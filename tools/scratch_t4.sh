python -m pytest tests/test_gpu_dense_nms.py -x -q > gpurun_out/dense_test.log 2>&1; tail -3 gpurun_out/dense_test.log
python -m pytest tests -m gpu -x -q -k "not config4 and not dense_nms" > gpurun_out/post_test.log 2>&1; tail -2 gpurun_out/post_test.log
echo cols64; python tools/bench_configs.py --only 5 2>&1 | tail -1 | cut -c1-130
echo cols128; B200VA_LIB=$PWD/realtime_video_analytics_32streams_b200/lib/libb200va_A.so python tools/bench_configs.py --only 5 2>&1 | tail -1 | cut -c1-130
echo cols128 ctas6; B200VA_DENSE_CTAS=6 B200VA_LIB=$PWD/realtime_video_analytics_32streams_b200/lib/libb200va_A.so python tools/bench_configs.py --only 5 2>&1 | tail -1 | cut -c1-130
B200VA_LIB=$PWD/realtime_video_analytics_32streams_b200/lib/libb200va_A.so python -m pytest tests/test_gpu_dense_nms.py -x -q 2>&1 | tail -1
