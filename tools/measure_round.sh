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
# the dense config (5): decode, k_dense_pairs, k_dense_resolve(_track), k_tracker
python tools/bench_configs.py --only 5 --steps 3 > $O/plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_dense|k_tracker$|k_decode' -s 8 -c 8 -f \
    -o $O/prof_${R}_dense python tools/bench_configs.py --only 5 --steps 3 > $O/ncu4.log 2>&1
ncu -i $O/prof_${R}_dense.ncu-rep --page raw --csv > $O/raw_${R}_dense.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_dense|k_tracker$|k_decode|k_letterbox' -c 400 --csv --log-file $O/${R}_launches_dense.csv \
    python tools/bench_configs.py --only 5 --steps 10 > $O/ncu5.log 2>&1
python tools/bench_configs.py --only 1,2,5,4,D,E --out $O/${R}_configs.json > $O/${R}_configs.log 2>&1
python tools/ncu_summary.py $O/raw_${R}.csv $O/${R}_launches.csv $O/${R}_ncu_summary.md "Round ${R#r} -- bench.py tick kernels (32 x 1080p + [32,84,8400]): ncu --set full and the launch list of the bench command" all
python tools/ncu_summary.py $O/raw_${R}_cfg4.csv $O/${R}_launches.csv $O/${R}_ncu_cfg4.md "Round ${R#r} -- config 4 (32 x 4K + ROI + motion): ncu --set full" all
python tools/ncu_summary.py $O/raw_${R}_dense.csv $O/${R}_launches_dense.csv $O/${R}_ncu_dense.md "Round ${R#r} -- config 5 (dense stress, 32 x ~1800 candidates): ncu --set full and the launch list" all
tail -c 600 $O/${R}_bench.json
