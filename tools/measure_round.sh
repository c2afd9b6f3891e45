#!/bin/bash
# Round measurement pass (run on the GPU box: `gpurun -- bash tools/measure_round.sh rN`).  Every ncu capture
# follows a plain run of the same command that exited 0; numbers printed under ncu are never bench values.
set -u
R=${1:-r2}
O=gpurun_out
mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/${R}_bench.json 2> $O/${R}_bench.err || { echo "bench failed"; tail -5 $O/${R}_bench.err; exit 1; }
python bench.py --impl reference --steps 20 --warmup 3 > $O/${R}_bench_reference.json 2>> $O/${R}_bench.err
python tools/bench_configs.py --only L --out $O/${R}_letterbox_shapes.json > $O/${R}_letterbox.log 2>&1
QUICK="--steps 20 --warmup 3 --no-cpu --no-configs --no-parity"
# launch list of the bench command (graph kernel nodes are listed individually)
python bench.py $QUICK > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${R}_launches.csv \
    python bench.py $QUICK > $O/ncu1.log 2>&1
# one full capture per kernel of the tick (eager launches so that every kernel is a plain launch)
python bench.py --steps 4 --warmup 3 --no-cpu --no-configs --no-parity --no-graph > $O/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_letterbox|k_decode|k_sort_nms|k_tracker$|k_post_track' -c 10 -f \
    -o $O/prof_${R} python bench.py --steps 4 --warmup 3 --no-cpu --no-configs --no-parity --no-graph > $O/ncu2.log 2>&1
ncu -i $O/prof_${R}.ncu-rep --page raw --csv > $O/raw_${R}.csv 2>/dev/null
# the motion tile kernel and the 4K + ROI letterbox (config 4)
python tools/bench_configs.py --only 4 --steps 3 > $O/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_motion|k_letterbox' -c 4 -f \
    -o $O/prof_${R}_cfg4 python tools/bench_configs.py --only 4 --steps 3 > $O/ncu3.log 2>&1
ncu -i $O/prof_${R}_cfg4.ncu-rep --page raw --csv > $O/raw_${R}_cfg4.csv 2>/dev/null
tail -c 600 $O/${R}_bench.json
