#!/bin/bash
# round 2, GPU call O: medium-band kernel: parity + A/B on C1, C2, C3 (both generators), C4
OUT=gpurun_out/r02o
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synth.py tests/test_gpu_power.py -m gpu -q --timeout 600 -p no:cacheprovider -x > $OUT/pytest_fast.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_fast.log; tail -12 $OUT/pytest_fast.log
for w in c1 c2; do
  for v in nobands mband auto; do
    timeout 120 python bench.py --workload $w --variant $v --steps 2000 --warmup 200 $B >> $OUT/small.jsonl 2>> $OUT/small.err
  done
  timeout 120 python bench.py --workload $w --variant mband --steps 200 --warmup 20 --cold $B >> $OUT/small.jsonl 2>> $OUT/small.err
done
for w in c3_spec c3 c4; do
  for v in nobands mband auto; do
    timeout 300 python bench.py --workload $w --variant $v --steps 20 --warmup 5 $B --breakdown >> $OUT/sweep.jsonl 2>> $OUT/sweep.err
  done
done
N="ncu --set full --clock-control none --import-source on"
cap() { name=$1; rx=$2; skip=$3; shift 3
  timeout 400 $N -k regex:$rx -s $skip -c 1 -f -o $OUT/$name "$@" > $OUT/ncu_$name.log 2>&1
  if [ -f $OUT/$name.ncu-rep ]; then
    python tools/ncu_summary.py $OUT/$name.ncu-rep > $OUT/$name.summary.txt 2>&1
    python tools/ncu_hot.py $OUT/$name.ncu-rep 25 > $OUT/$name.hot.txt 2>&1
    rm -f $OUT/$name.ncu-rep
  fi
}
cap c3spec_mb mb_kernel 3 python bench.py --workload c3_spec --variant mband $B --steps 3 --warmup 1
N="$N --cache-control none"
cap c1_mb mb_kernel 300 python bench.py --workload c1 --variant mband --steps 500 --warmup 100 $B
du -sh $OUT; echo done
