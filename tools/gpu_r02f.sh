#!/bin/bash
# round 2, GPU call F: short-band kernel with one 192 KB window per 16384-row band; LCB vs chunked on the sorted shapes
OUT=gpurun_out/r02f
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
for w in c5_spec c3_spec c5; do
  for v in auto nobands; do
    timeout 300 python bench.py --workload $w --variant $v --steps 20 --warmup 5 $B --breakdown >> $OUT/sweep.jsonl 2>> $OUT/sweep.err
  done
done
timeout 300 python bench.py --workload c5 --variant blocked --steps 20 --warmup 5 $B --breakdown >> $OUT/sweep.jsonl 2>> $OUT/sweep.err
timeout 300 python bench.py --workload c3 --variant blocked --steps 20 --warmup 5 $B --breakdown >> $OUT/sweep.jsonl 2>> $OUT/sweep.err
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synth.py -m gpu -q --timeout 600 -p no:cacheprovider -x > $OUT/pytest_fast.log 2>&1
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
cap c5spec_sb sb_kernel 3 python bench.py --workload c5_spec --steps 3 --warmup 1 $B
du -sh $OUT; echo done
