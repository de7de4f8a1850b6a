#!/bin/bash
OUT=gpurun_out/r02zi
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
run() { tag=$1; shift; echo "# $tag" >> $OUT/small.jsonl; timeout 120 env "$@" python bench.py --workload $W --steps 2000 --warmup 200 $B >> $OUT/small.jsonl 2>> $OUT/small.err; }
W=c1
run default X=1
run cta224 DASP_KEEP_CTA=224
run default_again X=1
run cta224_again DASP_KEEP_CTA=224
DASP_KEEP_CTA=224 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 500 -p no:cacheprovider -x -k "test_preprocessing_bit_exact_and_spmv and f64" > $OUT/pytest.log 2>&1
echo "pytest rc=$?"; tail -2 $OUT/pytest.log
echo done
