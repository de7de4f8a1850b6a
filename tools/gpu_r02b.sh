#!/bin/bash
# round 2, GPU call B: LCB v2, small-matrix CTA shapes, FP16 live tiles / HMMA, comparison harness, ncu captures
# (ncu reports are summarised ON THE BOX and deleted: gpurun_out may not exceed 64 MiB)
OUT=gpurun_out/r02b
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
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
for v in auto mma; do
  timeout 200 python bench.py --workload c4 --half --variant $v --steps 20 --warmup 5 $B >> $OUT/half.jsonl 2>> $OUT/half.err
  timeout 200 python bench.py --workload c2 --variant $v --steps 2000 --warmup 200 $B >> $OUT/half.jsonl 2>> $OUT/half.err
  for w in c3 c5; do timeout 200 python bench.py --workload ${w}_spec --half --variant $v --steps 20 --warmup 5 $B --breakdown >> $OUT/half.jsonl 2>> $OUT/half.err; done
done
for w in c1 c2 c3_spec c3 c4 c5_spec c5; do timeout 300 tools/compare $w >> $OUT/compare.jsonl 2>> $OUT/compare.err; done
# ncu: one full capture per kernel of interest, summarised here
N="ncu --set full --clock-control none --import-source on"
cap() { # name kernel-regex skip extra-ncu-args -- bench args
  name=$1; rx=$2; skip=$3; shift 3
  timeout 400 $N -k regex:$rx -s $skip -c 1 -f -o $OUT/$name "$@" > $OUT/ncu_$name.log 2>&1
  if [ -f $OUT/$name.ncu-rep ]; then
    python tools/ncu_summary.py $OUT/$name.ncu-rep > $OUT/$name.summary.txt 2>&1
    python tools/ncu_hot.py $OUT/$name.ncu-rep 30 > $OUT/$name.hot.txt 2>&1
    rm -f $OUT/$name.ncu-rep
  fi
}
cap c5spec_lcb lcb_kernel 3 python bench.py --workload c5_spec --steps 3 --warmup 1 $B
cap c3spec_lcb lcb_kernel 3 python bench.py --workload c3_spec --steps 3 --warmup 1 $B
cap c3spec_med spmv_kernel 3 python bench.py --workload c3_spec --categories 2 --steps 3 --warmup 1 $B
cap c5spec_short spmv_kernel 3 python bench.py --workload c5_spec --categories 4 --steps 3 --warmup 1 $B
N="$N --cache-control none"
cap c1_warm spmv_kernel 300 python bench.py --workload c1 --steps 500 --warmup 100 $B
cap c2_warm spmv_kernel 300 python bench.py --workload c2 --steps 500 --warmup 100 $B
du -sh $OUT
echo done
