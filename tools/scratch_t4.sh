python tools/bench_configs.py --only 1,2 > gpurun_out/cfg12.log 2>&1; tail -2 gpurun_out/cfg12.log | cut -c1-200
python bench.py --steps 20 --warmup 5 --no-cpu --no-configs > gpurun_out/bq.json 2>gpurun_out/bq.err; python -c "
import json; d=json.loads(open('gpurun_out/bq.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['kernel_ms'])"
