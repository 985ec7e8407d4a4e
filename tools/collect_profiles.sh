#!/bin/bash
# Run on the GPU box (via gpurun): ncu launch list of the bench command + one full capture per kernel.
# Every ncu run is preceded by the same command without ncu (B200_PROFILING.md).
set -u
OUT=gpurun_out/prof
mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > $OUT/bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches.csv $BENCH > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"
$BENCH > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lookup_fwd_px -s 300 -c 2 -o $OUT/lookup_cfg1 $BENCH > $OUT/ncu_lookup.log 2>&1
echo "lookup cfg1 rc=$?"
$BENCH > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:corr_build_kernel -s 8 -c 2 -o $OUT/build_cfg1 $BENCH > $OUT/ncu_build.log 2>&1
echo "build cfg1 rc=$?"
PROBE="python tools/lookup_probe.py cfg3"
$PROBE > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lookup_fwd_px -s 40 -c 2 -o $OUT/lookup_cfg3 $PROBE > $OUT/ncu_lookup3.log 2>&1
echo "lookup cfg3 rc=$?"
[ "${1:-}" = "--no-train" ] && { ls -la $OUT | head -30; exit 0; }
TRAIN="python tools/train_step_probe.py"
$TRAIN > $OUT/train_plain.log 2>&1 || exit 1
for k in lookup_bwd_batched lookup_bwd_acc fold_grad corr_build_bwd lookup_bwd_dense; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o $OUT/train_$k $TRAIN > $OUT/ncu_train_$k.log 2>&1
  echo "$k rc=$?"
done
ls -la $OUT | head -30
