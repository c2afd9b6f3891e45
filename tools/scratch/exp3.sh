# kernel durations for the small configs (1 and 4 streams): are the kernels themselves slow or is it launch gaps?
python tools/bench_configs.py --only 1,2 --steps 10 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/small_launches.csv python tools/bench_configs.py --only 1,2 --steps 10 > gpurun_out/small_ncu.log 2>&1
python - <<'P'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/small_launches.csv')) if len(r)>5]
h=rows[0]; k=h.index('Kernel Name'); v=h.index('Metric Value'); g=h.index('Grid Size') if 'Grid Size' in h else None
d=collections.OrderedDict()
for r in rows[1:]:
    name=r[k].split('(')[0].replace('<unnamed>::','').replace('void ','')
    if name.startswith('at::'): continue
    key=(name, r[g] if g is not None else '')
    d.setdefault(key,[]).append(float(r[v].replace(',',''))/1e3)
for (n,gr),x in d.items(): print(f"{n:28s} grid {gr:16s} n={len(x):3d} avg {sum(x)/len(x):6.1f} us  min {min(x):6.1f}")
P
