#!/bin/bash
OUT=gpurun_out/r02ze
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
run() { tag=$1; shift; echo "# $tag" >> $OUT/small.jsonl; timeout 120 env "$@" python bench.py --workload $W --steps 2000 --warmup 200 $B >> $OUT/small.jsonl 2>> $OUT/small.err; }
for W in c1 c2; do
  run default X=1
  run keep_order DASP_KEEP_ORDER=1
  run default_again X=1
  run keep_order_again DASP_KEEP_ORDER=1
done
N="ncu --set full --clock-control none --import-source on --cache-control none"
DASP_KEEP_ORDER=1 timeout 300 $N -k regex:spmv_kernel -s 300 -c 1 -f -o $OUT/c1_order python bench.py --workload c1 --steps 500 --warmup 100 $B > $OUT/ncu.log 2>&1
python tools/ncu_summary.py $OUT/c1_order.ncu-rep > $OUT/c1_order.summary.txt 2>&1; rm -f $OUT/c1_order.ncu-rep
echo done
