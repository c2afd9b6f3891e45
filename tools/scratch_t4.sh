python -m pytest tests -m gpu -x -q > gpurun_out/gputest.log 2>&1; tail -3 gpurun_out/gputest.log
python tools/bench_configs.py --only 1,2,5 2>&1 | tail -3 | cut -c1-150
