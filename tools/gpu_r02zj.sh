#!/bin/bash
# round 2, GPU call ZJ (same script as ZD, after the small-matrix changes): final validation of the shipped defaults: full GPU suite, smoke, both bench arms, launch lists, ncu of the shipped kernels
OUT=gpurun_out/r02zj
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider ) > $OUT/pytest_all.log 2>&1
tail -5 $OUT/pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > $OUT/bench_reference.json 2> $OUT/bench_reference.err
( time timeout 1200 python bench.py --steps 20 --warmup 5 ) > $OUT/bench_default.json 2> $OUT/bench_default.err
echo "bench rc=$?"; tail -4 $OUT/bench_default.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c4.csv python bench.py $B --steps 5 --warmup 3 > $OUT/launches_c4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_c5spec.csv python bench.py --workload c5_spec $B --steps 5 --warmup 3 > $OUT/launches_c5spec.log 2>&1
N="ncu --set full --clock-control none --import-source on"
cap() { name=$1; rx=$2; skip=$3; shift 3
  timeout 400 $N -k regex:$rx -s $skip -c 1 -f -o $OUT/$name "$@" > $OUT/ncu_$name.log 2>&1
  if [ -f $OUT/$name.ncu-rep ]; then
    python tools/ncu_summary.py $OUT/$name.ncu-rep > $OUT/$name.summary.txt 2>&1
    python tools/ncu_hot.py $OUT/$name.ncu-rep 25 > $OUT/$name.hot.txt 2>&1
    rm -f $OUT/$name.ncu-rep
  fi
}
cap c4_fused spmv_kernel 3 python bench.py $B --steps 3 --warmup 1
cap c5spec_lcb lcb_kernel 3 python bench.py --workload c5_spec $B --steps 3 --warmup 1
cap c3spec_lcb lcb_kernel 3 python bench.py --workload c3_spec $B --steps 3 --warmup 1
N="$N --cache-control none"
cap c1_fused spmv_kernel 300 python bench.py --workload c1 --steps 500 --warmup 100 $B
cap c2_fused spmv_kernel 300 python bench.py --workload c2 --steps 500 --warmup 100 $B
for w in c1 c2 c3_spec c4 c5_spec; do timeout 300 tools/compare $w >> $OUT/compare.jsonl 2>> $OUT/compare.err; done
du -sh $OUT; echo done
