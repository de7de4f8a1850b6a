#!/bin/bash
# round 2, GPU call R: SM-affine queues v2 (8-group chunks, k = resident CTAs per SM, overflow by block index), lean variant; slab pools
OUT=gpurun_out/r02r
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synth.py tests/test_gpu_power.py -m gpu -q --timeout 600 -p no:cacheprovider -x > $OUT/pytest_fast.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_fast.log; tail -5 $OUT/pytest_fast.log
run() { tag=$1; shift; echo "# $tag" >> $OUT/small.jsonl; timeout 120 env "$@" python bench.py --workload $W --steps 2000 --warmup 200 $B >> $OUT/small.jsonl 2>> $OUT/small.err; }
for W in c1 c2; do
  run nosmq DASP_NO_SMQ=1
  run smq_pipe X=1
  run smq_lean DASP_SMQ_LEAN=1
  run smq_pipe_again X=1
  echo "# cold smq_pipe" >> $OUT/small.jsonl
  timeout 120 python bench.py --workload $W --steps 200 --warmup 20 --cold $B >> $OUT/small.jsonl 2>> $OUT/small.err
done
echo "# c4 preprocess" >> $OUT/small.jsonl
timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 $B >> $OUT/small.jsonl 2>> $OUT/small.err
N="ncu --set full --clock-control none --import-source on --cache-control none"
cap() { name=$1; rx=$2; skip=$3; shift 3
  timeout 400 $N -k regex:$rx -s $skip -c 1 -f -o $OUT/$name "$@" > $OUT/ncu_$name.log 2>&1
  if [ -f $OUT/$name.ncu-rep ]; then
    python tools/ncu_summary.py $OUT/$name.ncu-rep > $OUT/$name.summary.txt 2>&1
    python tools/ncu_hot.py $OUT/$name.ncu-rep 25 > $OUT/$name.hot.txt 2>&1
    rm -f $OUT/$name.ncu-rep
  fi
}
cap c1_smq smq_kernel 300 python bench.py --workload c1 --steps 500 --warmup 100 $B
DASP_SMQ_LEAN=1 cap c1_smq_lean smq_kernel 300 python bench.py --workload c1 --steps 500 --warmup 100 $B
cap c2_smq smq_kernel 300 python bench.py --workload c2 --steps 500 --warmup 100 $B
du -sh $OUT; echo done
