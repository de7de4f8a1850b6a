#!/bin/bash
# round 2, GPU call ZK: memcheck / racecheck on the final defaults (small-matrix lean + order path, 224-thread CTAs), preprocessing times in a fresh process
OUT=gpurun_out/r02zk
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
B="--no-secondary --no-cpu --no-others --no-iterated"
( time timeout 1200 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_power.py tests/test_gpu_synth.py -m gpu -q -x --timeout 1100 -p no:cacheprovider -k "not reference_main and not c_example and not sm_affine" ) > $OUT/memcheck.log 2>&1
echo "memcheck rc=$?" >> $OUT/memcheck.log; grep -E "passed|ERROR SUMMARY|rc=" $OUT/memcheck.log
( time timeout 600 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 500 -p no:cacheprovider -k "test_preprocessing_bit_exact_and_spmv and (skewed or powerlaw or mixed or one_row or long_pad or lcb_row)" ) > $OUT/racecheck.log 2>&1
echo "racecheck rc=$?" >> $OUT/racecheck.log; grep -E "passed|RACECHECK SUMMARY|rc=" $OUT/racecheck.log
for w in c3_spec c5_spec; do timeout 300 python bench.py --workload $w --steps 20 --warmup 5 $B >> $OUT/pre.jsonl 2>> $OUT/pre.err; done
python - <<'P'
import json
for l in open('gpurun_out/r02zk/pre.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['workload'][:12], d['ms_per_step'], d['preprocess']['gpu_ms'], d['preprocess']['first_create_in_process_gpu_ms'])
P
echo done
