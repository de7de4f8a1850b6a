#!/bin/bash
# round 2, GPU call T: dasp_create timing, one process per allocator mode (first create = what bench.py reports)
OUT=gpurun_out/r02t
mkdir -p $OUT
for mode in "DASP_NO_SLAB=1" "X=1" "DASP_SLAB_MAX_MB=1024" "DASP_NO_SLAB=1" "X=1"; do
  timeout 300 python tools/preprocess_time.py 256 $mode >> $OUT/preprocess_time.txt 2>&1
done
cat $OUT/preprocess_time.txt
