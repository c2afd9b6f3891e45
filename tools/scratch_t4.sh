python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_tick.py tests/test_gpu_ultralytics.py tests/test_gpu_properties.py tests/test_collector.py tests/test_gpu_baseline_shapes.py tests/test_c_consumer.py -x -q -k "not config4" > gpurun_out/post_test.log 2>&1; tail -3 gpurun_out/post_test.log
for i in 1 2; do
echo round-start; (cd tools/scratch/R && python tools/bench_configs.py --only 1,2 2>&1 | tail -2 | cut -c1-120)
echo now; python tools/bench_configs.py --only 1,2,5 2>&1 | tail -3 | cut -c1-120
done
ncu --set full --clock-control none --import-source on -k regex:k_post_track -s 6 -c 4 -f -o gpurun_out/prof_small python tools/scratch_small.py > gpurun_out/ncu_small.log 2>&1; tail -2 gpurun_out/ncu_small.log
