#!/bin/bash
# ncu launch list of the headline bench command (BASELINE config 1 only): per-launch gpu__time_duration
# of every kernel the command launches -- preceded by the same command without ncu.
set -u
OUT=gpurun_out/prof2
mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 3 --reps 1 --no-cpu-baseline --configs none"
$BENCH > $OUT/bench_plain_headline.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv $BENCH > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$? ($(grep -c '^"' $OUT/launches.csv) rows)"
