#!/bin/bash
# round 2, GPU call Y: small matrices: compact column indices in the pipelined loop vs reg_cid vs the register-lean loop
OUT=gpurun_out/r02y
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
DASP_KEEP_COMPACT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 500 -p no:cacheprovider -x -k "test_preprocessing_bit_exact_and_spmv or fuzz" > $OUT/pytest_compact.log 2>&1
echo "compact pytest rc=$?"; tail -2 $OUT/pytest_compact.log
run() { tag=$1; shift; echo "# $tag" >> $OUT/small.jsonl; timeout 120 env "$@" python bench.py --workload $W --steps 2000 --warmup 200 $B >> $OUT/small.jsonl 2>> $OUT/small.err; }
for W in c1 c2; do
  run shipped X=1
  run pipe_compact DASP_KEEP_COMPACT=1
  run keep_lean DASP_KEEP_LEAN=1
  run pipe_compact_again DASP_KEEP_COMPACT=1
  echo "# cold pipe_compact" >> $OUT/small.jsonl
  DASP_KEEP_COMPACT=1 timeout 120 python bench.py --workload $W --steps 200 --warmup 20 --cold $B >> $OUT/small.jsonl 2>> $OUT/small.err
  echo "# cold keep_lean" >> $OUT/small.jsonl
  DASP_KEEP_LEAN=1 timeout 120 python bench.py --workload $W --steps 200 --warmup 20 --cold $B >> $OUT/small.jsonl 2>> $OUT/small.err
done
echo done
