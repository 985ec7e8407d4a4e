#!/bin/bash
# thread-per-pixel lookup kernel: pixels per CTA x (one thread per pixel | per level pair)
for split in ${SPLITS:-0 1}; do for tpb in ${TPBS:-32 64 128}; do
  echo "== SPLIT=$split TPB=$tpb"
  SMOE_CORR_PX_SPLIT=$split SMOE_CORR_PX_TPB=$tpb python tools/lookup_kernels_probe.py "$@" 2>&1 | sed -E 's/vec [^,]*, pair [^,]*, //'
done; done
