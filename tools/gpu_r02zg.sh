#!/bin/bash
OUT=gpurun_out/r02zg
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
run() { tag=$1; shift; echo "# $tag" >> $OUT/small.jsonl; timeout 120 env "$@" python bench.py --workload $W --steps 2000 --warmup 200 $B >> $OUT/small.jsonl 2>> $OUT/small.err; }
for W in c1 c2; do
  run order_lean_tb4 DASP_KEEP_ORDER=1 DASP_KEEP_LEAN=1
  run order_lean_tb3 DASP_KEEP_ORDER=1 DASP_KEEP_LEAN=6
  run order_lean_tb2 DASP_KEEP_ORDER=1 DASP_KEEP_LEAN=7
done
N="ncu --set full --clock-control none --import-source on --cache-control none"
DASP_KEEP_ORDER=1 DASP_KEEP_LEAN=1 timeout 300 $N -k regex:spmv_kernel -s 300 -c 1 -f -o $OUT/c1_order_lean python bench.py --workload c1 --steps 500 --warmup 100 $B > $OUT/ncu.log 2>&1
python tools/ncu_summary.py $OUT/c1_order_lean.ncu-rep > $OUT/c1_order_lean.summary.txt 2>&1; python tools/ncu_hot.py $OUT/c1_order_lean.ncu-rep 20 > $OUT/c1_order_lean.hot.txt 2>&1; rm -f $OUT/c1_order_lean.ncu-rep
DASP_KEEP_ORDER=1 timeout 300 $N -k regex:spmv_kernel -s 300 -c 1 -f -o $OUT/c2_order_lean python bench.py --workload c2 --steps 500 --warmup 100 $B > $OUT/ncu2.log 2>&1
python tools/ncu_summary.py $OUT/c2_order_lean.ncu-rep > $OUT/c2_order_lean.summary.txt 2>&1; rm -f $OUT/c2_order_lean.ncu-rep
echo done
