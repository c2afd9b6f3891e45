python -m pytest tests -m gpu -x -q -k "not config4" > gpurun_out/post_test.log 2>&1; tail -2 gpurun_out/post_test.log
echo new; python tools/bench_configs.py --only 5 2>&1 | tail -1 | cut -c1-200
echo old; B200VA_DENSE_IMPL=1 python tools/bench_configs.py --only 5 2>&1 | tail -1 | cut -c1-200
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_dense -c 3 --csv python tools/bench_configs.py --only 5 --steps 3 2>/dev/null | grep -E "k_dense" | cut -d, -f5,15-
