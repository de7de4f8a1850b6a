#!/bin/bash
# round 2, GPU call W: column-blocked long rows BESIDE the fused kernel (side stream), C3 both generators; parity first
OUT=gpurun_out/r02w
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
for ov in 3 2; do
  DASP_LCB_OVERLAP=$ov timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synth.py -m gpu -q --timeout 500 -p no:cacheprovider -x -k "not sm_affine and not reference_main and not c_example" > $OUT/pytest_overlap$ov.log 2>&1
  echo "overlap=$ov pytest rc=$?"; tail -2 $OUT/pytest_overlap$ov.log
done
run() { tag=$1; shift; echo "# $tag" >> $OUT/sweep.jsonl; timeout 300 env "$@" python bench.py --workload $W --steps 30 --warmup 5 $B >> $OUT/sweep.jsonl 2>> $OUT/sweep.err; }
for W in c3_spec c3; do
  run serial X=1
  run overlap3 DASP_LCB_OVERLAP=3
  run overlap2 DASP_LCB_OVERLAP=2
  run overlap1 DASP_LCB_OVERLAP=1
done
W=c3_spec
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_c3spec_overlap2.csv env DASP_LCB_OVERLAP=2 python bench.py --workload c3_spec $B --steps 3 --warmup 2 > $OUT/launches.log 2>&1
echo done
