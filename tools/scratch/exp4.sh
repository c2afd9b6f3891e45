cd realtime_video_analytics_32streams_b200/csrc
for mb in 6 7 8; do
  rm -f preprocess.o; make NVFLAGS_EXTRA="-DLB_MINBLOCKS=$mb" -j8 > /dev/null 2>&1 || { echo build failed; exit 1; }
  grep -A2 "k_letterboxILi0ELb0" preprocess.ptxas.log | grep -i "registers" | head -1
  cd ../..
  echo "minblocks $mb"; python tools/bench_configs.py --only L 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('  ', d['config'], d['preprocess'], round(d['frac_of_peak'],3))
"
  python bench.py --no-cpu --steps 100 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  bench', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])
"
  cd realtime_video_analytics_32streams_b200/csrc
done
