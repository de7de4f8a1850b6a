#!/bin/bash
# round 2, GPU call Z: FP64 small matrices: pipelined loop with one tile per batch at 64 registers (4 CTAs per SM, one wave); defaults
OUT=gpurun_out/r02z
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
DASP_KEEP_LEAN=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 500 -p no:cacheprovider -x -k "test_preprocessing_bit_exact_and_spmv and f64" > $OUT/pytest_b1.log 2>&1
echo "b1 pytest rc=$?"; tail -2 $OUT/pytest_b1.log
run() { tag=$1; shift; echo "# $tag" >> $OUT/small.jsonl; timeout 120 env "$@" python bench.py --workload $W --steps 2000 --warmup 200 $B >> $OUT/small.jsonl 2>> $OUT/small.err; }
W=c1
run default X=1
run pipe_b1_4ctas DASP_KEEP_LEAN=2
run default_again X=1
run pipe_b1_4ctas_again DASP_KEEP_LEAN=2
W=c2
run default X=1
run f16_pipelined DASP_KEEP_LEAN=0
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_power.py -m gpu -q --timeout 500 -p no:cacheprovider -x > $OUT/pytest_default.log 2>&1
echo "default pytest rc=$?"; tail -2 $OUT/pytest_default.log
echo done
