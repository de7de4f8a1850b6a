#!/bin/bash
# round 2, GPU call ZL: the small-matrix AUTO choice (lean loop + locality order) against the round-1 shape on OTHER small matrices
OUT=gpurun_out/r02zl
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
run() { tag=$1; shift; echo "# $tag" >> $OUT/small.jsonl; timeout 120 env "$@" python bench.py "${WL[@]}" --steps 1000 --warmup 100 $B >> $OUT/small.jsonl 2>> $OUT/small.err; }
WL=(--workload c4 --grid 40); run "stencil40 auto" X=1; run "stencil40 round1_shape" DASP_KEEP_LEAN=0
WL=(--workload c4 --grid 40 --half); run "stencil40_f16 auto" X=1; run "stencil40_f16 round1_shape" DASP_KEEP_LEAN=0
WL=(--workload c3_spec --scale 0.012); run "powerlaw120k auto" X=1; run "powerlaw120k round1_shape" DASP_KEEP_LEAN=0
WL=(--workload c3_spec --scale 0.012 --half); run "powerlaw120k_f16 auto" X=1; run "powerlaw120k_f16 round1_shape" DASP_KEEP_LEAN=0
WL=(--workload c5_spec --scale 0.004); run "skewed200k auto" X=1; run "skewed200k round1_shape" DASP_KEEP_LEAN=0
WL=(--workload c1); run "c1 auto(tb3,224)" X=1; run "c1 tb4_224" DASP_KEEP_LEAN=1; run "c1 tb3_256" DASP_KEEP_CTA=256
tail -3 $OUT/small.err
echo done
