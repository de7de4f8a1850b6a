#!/bin/bash
OUT=gpurun_out/r02za
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
run() { tag=$1; shift; echo "# $tag" >> $OUT/small.jsonl; timeout 120 env "$@" python bench.py --workload $W --steps 2000 --warmup 200 $B >> $OUT/small.jsonl 2>> $OUT/small.err; }
W=c1
run default X=1
run pipe_3ctas DASP_KEEP_LEAN=3
run default_again X=1
run pipe_3ctas_again DASP_KEEP_LEAN=3
echo done
