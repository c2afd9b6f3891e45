for i in 1 2; do
echo round-start; (cd tools/scratch/R && python tools/bench_configs.py --only 1,2 2>&1 | tail -2 | cut -c1-120)
echo now; python tools/bench_configs.py --only 1,2 2>&1 | tail -2 | cut -c1-120
done
