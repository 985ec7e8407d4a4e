#!/bin/bash
# Run on the GPU box (via gpurun): the ncu evidence of round 2 at HEAD -- launch list of the bench
# command and one `--set full` capture per (kernel class, BASELINE config) the bench line reports a
# roofline for.  Every ncu run is preceded by the same command without ncu (B200_PROFILING.md).
set -u
OUT=gpurun_out/prof2
mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs cfg2,cfg3"
$BENCH > $OUT/bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/launches.csv $BENCH > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"
cap() {   # cfg class kernel-regex skip
  local T="python tools/ncu_target.py $1 $2"
  $T > $OUT/plain_$1_$2.log 2>&1 || { echo "$1 $2 plain run failed"; return; }
  ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -o $OUT/$2_$1 $T > $OUT/ncu_$1_$2.log 2>&1
  echo "$1 $2 rc=$? $(cat $OUT/plain_$1_$2.log | tail -1)"
}
cap cfg1 build corr_build_kernel 2
cap cfg2 build corr_build_kernel 2
cap cfg3 build corr_build_kernel 2
cap cfg1 lookup lookup_fwd 2
cap cfg2 lookup lookup_fwd 2
cap cfg3 lookup lookup_fwd 2
cap cfg4 lookup lookup_fwd 2
cap cfg2 lookup_bwd lookup_bwd 2
cap cfg2 build_bwd corr_build_bwd 1
ls -la $OUT | head -40
