python -m pytest tests -m gpu -x -q -k "not config4" > gpurun_out/post_test.log 2>&1; tail -3 gpurun_out/post_test.log
python tools/bench_configs.py --only 5 2>&1 | tail -1 | cut -c1-150
B200VA_BENCH_SCHEDULE=4 python tools/bench_configs.py --only 5 2>&1 | tail -1 | cut -c1-150
B200VA_BENCH_SCHEDULE=3 python tools/bench_configs.py --only 5 2>&1 | tail -1 | cut -c1-150
