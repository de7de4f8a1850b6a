#!/bin/bash
# round 2, GPU call ZP: adaptive long-row work units + minimum size for the column-blocked kernel: parity, small long-row matrices, C3 sanity
OUT=gpurun_out/r02zp
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_power.py tests/test_gpu_synth.py -m gpu -q --timeout 500 -p no:cacheprovider -x > $OUT/pytest.log 2>&1
echo "pytest rc=$?"; tail -2 $OUT/pytest.log
run() { tag=$1; shift; echo "# $tag" >> $OUT/small.jsonl; timeout 120 env "$@" python bench.py "${WL[@]}" --steps $S --warmup 20 $B --breakdown >> $OUT/small.jsonl 2>> $OUT/small.err; }
S=1000
WL=(--workload c3_spec --scale 0.012); run "powerlaw120k auto" X=1; run "powerlaw120k units32" DASP_LONG_UNIT_WARPS=32
WL=(--workload c5_spec --scale 0.004); run "skewed200k auto" X=1
WL=(--workload c3_spec --scale 0.05); run "powerlaw500k auto" X=1
S=20
WL=(--workload c3_spec); run "c3_spec auto" X=1
tail -3 $OUT/small.err
echo done
