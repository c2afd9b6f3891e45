python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for b in 32768 57344 81920 204800; do echo "budget $b"; B200VA_LB_SMEM_BUDGET=$b python tools/bench_configs.py --only L 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('  ', d['config'], d['preprocess'], round(d['frac_of_peak'],3))
"; done
for m in 0 4 8; do echo "split $m"; B200VA_DECODE_SPLIT=$m python bench.py --no-cpu --steps 100 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  ', d['value'], d['ms_per_step'], d['breakdown_ms'], d['roofline']['kernel_ms'])
"; done
