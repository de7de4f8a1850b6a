#!/bin/bash
# round 2, GPU call S: dasp_create timing (slab allocator A/B + phase trace), parity after the allocator change, C1/C2 defaults
OUT=gpurun_out/r02s
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
timeout 600 python tools/preprocess_time.py 256 > $OUT/preprocess_time.txt 2>&1; tail -30 $OUT/preprocess_time.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synth.py tests/test_gpu_power.py -m gpu -q --timeout 600 -p no:cacheprovider -x > $OUT/pytest_fast.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_fast.log; tail -3 $OUT/pytest_fast.log
for W in c1 c2; do timeout 120 python bench.py --workload $W --steps 2000 --warmup 200 $B >> $OUT/small.jsonl 2>> $OUT/small.err; done
timeout 300 python bench.py --workload c3_spec --steps 20 --warmup 5 $B >> $OUT/small.jsonl 2>> $OUT/small.err
timeout 300 python bench.py --workload c5_spec --steps 20 --warmup 5 $B >> $OUT/small.jsonl 2>> $OUT/small.err
echo done
