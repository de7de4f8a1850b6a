#!/bin/bash
OUT=gpurun_out/r02zh
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
run() { tag=$1; shift; echo "# $tag" >> $OUT/small.jsonl; timeout 120 env "$@" python bench.py --workload $W --steps 2000 --warmup 200 $B >> $OUT/small.jsonl 2>> $OUT/small.err; }
for W in c1 c2; do
  run default_lean_order_irp X=1
  run cta128 DASP_KEEP_CTA=128
  run default_again X=1
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_power.py -m gpu -q --timeout 500 -p no:cacheprovider -x > $OUT/pytest.log 2>&1
echo "pytest rc=$?"; tail -2 $OUT/pytest.log
echo done
