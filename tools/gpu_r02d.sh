#!/bin/bash
# round 2, GPU call D: locality-ordered work lists (A/B), LCB 3-stage pipeline, C4 regression, tests
OUT=gpurun_out/r02d
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
for w in c3_spec c3 c5_spec c5 c4; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 $B --breakdown | sed 's/^{/{"order": 1, /' >> $OUT/sweep.jsonl 2>> $OUT/sweep.err
  DASP_NO_LOCALITY_ORDER=1 timeout 300 python bench.py --workload $w --steps 20 --warmup 5 $B --breakdown | sed 's/^{/{"order": 0, /' >> $OUT/sweep.jsonl 2>> $OUT/sweep.err
done
timeout 300 python bench.py --workload c3 --variant cuda --steps 20 --warmup 5 $B --breakdown | sed 's/^{/{"order": 1, /' >> $OUT/sweep.jsonl 2>> $OUT/sweep.err
timeout 300 python bench.py --workload c4 --half --steps 20 --warmup 5 $B | sed 's/^{/{"order": 1, /' >> $OUT/sweep.jsonl 2>> $OUT/sweep.err
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synth.py tests/test_gpu_power.py tests/test_gpu_multi.py -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/pytest_fast.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_fast.log; tail -3 $OUT/pytest_fast.log
N="ncu --set full --clock-control none --import-source on"
cap() { name=$1; rx=$2; skip=$3; shift 3
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
du -sh $OUT; echo done
