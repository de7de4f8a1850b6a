#!/bin/bash
# round 2, GPU call B: LCB v2, small-matrix CTA shapes, FP16 live tiles, comparison harness, ncu captures
OUT=gpurun_out/r02b
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synth.py tests/test_gpu_power.py -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/pytest_fast.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_fast.log; tail -3 $OUT/pytest_fast.log
for w in c3_spec c3 c5_spec c5; do
  for v in auto cuda blocked; do
    timeout 300 python bench.py --workload $w --variant $v --steps 20 --warmup 5 $B --breakdown >> $OUT/sweep.jsonl 2>> $OUT/sweep.err
  done
done
for shape in 128 256; do
  for w in c1 c2; do
    DASP_KEEP_CTA=$shape timeout 120 python bench.py --workload $w --steps 2000 --warmup 200 $B | sed "s/^{/{\"keep_cta\": $shape, /" >> $OUT/small.jsonl 2>> $OUT/small.err
    DASP_KEEP_CTA=$shape timeout 120 python bench.py --workload $w --steps 200 --warmup 20 --cold $B | sed "s/^{/{\"keep_cta\": $shape, /" >> $OUT/small.jsonl 2>> $OUT/small.err
  done
done
timeout 200 python bench.py --workload c4 --half --steps 20 --warmup 5 $B >> $OUT/half.jsonl 2>> $OUT/half.err
for w in c3 c5; do timeout 200 python bench.py --workload ${w}_spec --half --steps 20 --warmup 5 $B >> $OUT/half.jsonl 2>> $OUT/half.err; done
for w in c1 c2 c3_spec c3 c4 c5_spec c5; do timeout 300 tools/compare $w >> $OUT/compare.jsonl 2>> $OUT/compare.err; done
# ncu: one full capture per kernel of interest
N="ncu --set full --clock-control none --import-source on"
timeout 300 $N -k regex:lcb_kernel -s 3 -c 1 -o $OUT/prof_c5spec_lcb python bench.py --workload c5_spec --steps 3 --warmup 1 $B > $OUT/ncu_c5spec_lcb.log 2>&1
timeout 300 $N -k regex:lcb_kernel -s 3 -c 1 -o $OUT/prof_c3spec_lcb python bench.py --workload c3_spec --steps 3 --warmup 1 $B > $OUT/ncu_c3spec_lcb.log 2>&1
timeout 300 $N -k regex:spmv_kernel -s 3 -c 1 -o $OUT/prof_c3spec_med python bench.py --workload c3_spec --categories 2 --steps 3 --warmup 1 $B > $OUT/ncu_c3spec_med.log 2>&1
timeout 300 $N -k regex:spmv_kernel -s 3 -c 1 -o $OUT/prof_c5spec_short python bench.py --workload c5_spec --categories 4 --steps 3 --warmup 1 $B > $OUT/ncu_c5spec_short.log 2>&1
timeout 300 $N --cache-control none -k regex:spmv_kernel -s 300 -c 1 -o $OUT/prof_c1_warm python bench.py --workload c1 --steps 500 --warmup 100 $B > $OUT/ncu_c1.log 2>&1
timeout 300 $N --cache-control none -k regex:spmv_kernel -s 300 -c 1 -o $OUT/prof_c2_warm python bench.py --workload c2 --steps 500 --warmup 100 $B > $OUT/ncu_c2.log 2>&1
ls -la $OUT
echo done
