#!/bin/bash
# round 2, GPU call H: why is a half slab of C5 slow? (rank 0 of 2 on one GPU, per-category breakdown), SB window fix
OUT=gpurun_out/r02h
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
timeout 300 python bench.py --workload c5_spec --slab 0/2 --steps 20 --warmup 5 $B --breakdown >> $OUT/slab.jsonl 2>> $OUT/slab.err
timeout 300 python bench.py --workload c5_spec --slab 0/8 --steps 20 --warmup 5 $B --breakdown >> $OUT/slab.jsonl 2>> $OUT/slab.err
timeout 300 python bench.py --workload c5_spec --steps 20 --warmup 5 $B --breakdown >> $OUT/slab.jsonl 2>> $OUT/slab.err
timeout 300 python bench.py --workload c4 --slab 0/8 --steps 50 --warmup 5 $B >> $OUT/slab.jsonl 2>> $OUT/slab.err
N="ncu --set full --clock-control none --import-source on"
cap() { name=$1; rx=$2; skip=$3; shift 3
  timeout 400 $N -k regex:$rx -s $skip -c 1 -f -o $OUT/$name "$@" > $OUT/ncu_$name.log 2>&1
  if [ -f $OUT/$name.ncu-rep ]; then
    python tools/ncu_summary.py $OUT/$name.ncu-rep > $OUT/$name.summary.txt 2>&1
    python tools/ncu_hot.py $OUT/$name.ncu-rep 30 > $OUT/$name.hot.txt 2>&1
    rm -f $OUT/$name.ncu-rep
  fi
}
cap slab2_lcb lcb_kernel 3 python bench.py --workload c5_spec --slab 0/2 --steps 3 --warmup 1 $B
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synth.py -m gpu -q --timeout 600 -p no:cacheprovider -x > $OUT/pytest_fast.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_fast.log; tail -3 $OUT/pytest_fast.log
echo done
