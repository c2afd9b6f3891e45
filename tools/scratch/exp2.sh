for t in 256 512 1024; do echo "trk threads $t"; B200VA_TRK_THREADS=$t python -m pytest tests -m gpu -x -q -k "tracker or tick or pipeline or engine" 2>&1 | tail -1; B200VA_TRK_THREADS=$t python tools/bench_configs.py --only 2,5 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('  ', d['config'][:30], 'trk', d['tracker'], 'post', d['postprocess'])
"; B200VA_TRK_THREADS=$t python bench.py --no-cpu --steps 100 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  bench', d['value'], d['ms_per_step'], d['breakdown_ms'])
"; done
