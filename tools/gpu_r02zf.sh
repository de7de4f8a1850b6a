#!/bin/bash
OUT=gpurun_out/r02zf
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
run() { tag=$1; shift; echo "# $tag" >> $OUT/small.jsonl; timeout 120 env "$@" python bench.py --workload $W --steps 2000 --warmup 200 $B >> $OUT/small.jsonl 2>> $OUT/small.err; }
W=c2
run f16_order_lean DASP_KEEP_ORDER=1
run f16_order_pipe DASP_KEEP_ORDER=1 DASP_KEEP_LEAN=0
run f16_order_lean_smq DASP_KEEP_ORDER=1 DASP_SMQ=1 DASP_SMQ_LEAN=1
W=c1
run f64_order_lean DASP_KEEP_ORDER=1 DASP_KEEP_LEAN=1
run f64_order_pipe3 DASP_KEEP_ORDER=1 DASP_KEEP_LEAN=3
echo "# cold f16_order_lean" >> $OUT/small.jsonl
W=c2; DASP_KEEP_ORDER=1 timeout 120 python bench.py --workload $W --steps 200 --warmup 20 --cold $B >> $OUT/small.jsonl 2>> $OUT/small.err
DASP_KEEP_ORDER=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 500 -p no:cacheprovider -x -k "test_preprocessing_bit_exact_and_spmv or fuzz" > $OUT/pytest_order.log 2>&1
echo "order pytest rc=$?"; tail -2 $OUT/pytest_order.log
echo done
