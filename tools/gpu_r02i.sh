#!/bin/bash
OUT=gpurun_out/r02i
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
for s in 0/2 0/8; do timeout 300 python bench.py --workload c5_spec --slab $s --steps 20 --warmup 5 $B --breakdown >> $OUT/slab.jsonl 2>> $OUT/slab.err; done
timeout 300 python bench.py --workload c5_spec --steps 20 --warmup 5 $B --breakdown >> $OUT/slab.jsonl 2>> $OUT/slab.err
timeout 300 python bench.py --workload c3_spec --steps 20 --warmup 5 $B --breakdown >> $OUT/slab.jsonl 2>> $OUT/slab.err
timeout 300 python bench.py --workload c5_spec --slab 0/8 --variant banded --steps 20 --warmup 5 $B --breakdown >> $OUT/slab.jsonl 2>> $OUT/slab.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "preprocessing_bit_exact" > $OUT/pytest_fast.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_fast.log; tail -3 $OUT/pytest_fast.log
echo done
