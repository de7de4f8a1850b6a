#!/bin/bash
# round 2, GPU call X: small matrices through the register-lean loop (64 registers, 4 CTAs per SM, one wave) vs the shipped shape
OUT=gpurun_out/r02x
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
DASP_KEEP_LEAN=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 500 -p no:cacheprovider -x -k "test_preprocessing_bit_exact_and_spmv or fuzz" > $OUT/pytest_lean.log 2>&1
echo "lean pytest rc=$?"; tail -2 $OUT/pytest_lean.log
run() { tag=$1; shift; echo "# $tag" >> $OUT/small.jsonl; timeout 120 env "$@" python bench.py --workload $W --steps 2000 --warmup 200 $B >> $OUT/small.jsonl 2>> $OUT/small.err; }
for W in c1 c2; do
  run shipped X=1
  run keep_lean DASP_KEEP_LEAN=1
  run shipped_again X=1
  run keep_lean_again DASP_KEEP_LEAN=1
done
echo done
