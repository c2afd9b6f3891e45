for f in 0 33792 40960 49152; do echo floor $f; B200VA_LB_SMEM_FLOOR=$f python tools/bench_configs.py --only L --lshape 0 2>&1 | tail -1 | cut -c1-140; done
echo; for f in 0 33792; do echo floor $f all shapes; B200VA_LB_SMEM_FLOOR=$f python tools/bench_configs.py --only L 2>&1 | tail -6 | cut -c1-60,100-180; done
