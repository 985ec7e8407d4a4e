#!/bin/bash
set -u
OUT=gpurun_out/prof
mkdir -p $OUT
TRAIN="python tools/train_step_probe.py"
$TRAIN > $OUT/train_plain.log 2>&1 || exit 1
for k in lookup_bwd_acc fold_grad corr_build_bwd lookup_bwd_dense; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o $OUT/train_$k $TRAIN > $OUT/ncu_train_$k.log 2>&1
  echo "$k rc=$?"
done
