#!/bin/bash
# round 2, GPU call ZC: bank-aware order inside the runs of the column-blocked long rows: parity, A/B on C5 / C3, with and without the 16-bit indices
OUT=gpurun_out/r02zc
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synth.py tests/test_gpu_power.py -m gpu -q --timeout 600 -p no:cacheprovider -x > $OUT/pytest_fast.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_fast.log; tail -4 $OUT/pytest_fast.log
DASP_LCB_IDX16=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "test_preprocessing_bit_exact_and_spmv or fuzz" > $OUT/pytest_idx16.log 2>&1
echo "pytest idx16 rc=$?" >> $OUT/pytest_idx16.log; tail -3 $OUT/pytest_idx16.log
run() { tag=$1; shift; echo "# $tag" >> $OUT/sweep.jsonl; timeout 400 env "$@" python bench.py --workload $W --steps 20 --warmup 5 $B --breakdown >> $OUT/sweep.jsonl 2>> $OUT/sweep.err; }
for W in c5_spec c3_spec; do
  run csr_order DASP_LCB_BANK_ORDER=0
  run bank_order X=1
  run bank_order_idx16 DASP_LCB_IDX16=1
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
cap c5spec_lcb_bank lcb_kernel 3 python bench.py --workload c5_spec $B --steps 3 --warmup 1
DASP_LCB_IDX16=1 cap c5spec_lcb_bank16 lcb_kernel 3 python bench.py --workload c5_spec $B --steps 3 --warmup 1
tail -3 $OUT/sweep.err
echo done
