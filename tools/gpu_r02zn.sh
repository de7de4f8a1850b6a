#!/bin/bash
OUT=gpurun_out/r02zn
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
run() { tag=$1; shift; echo "# $tag" >> $OUT/small.jsonl; timeout 120 env "$@" python bench.py "${WL[@]}" --steps 1000 --warmup 100 $B --breakdown >> $OUT/small.jsonl 2>> $OUT/small.err; }
WL=(--workload c3_spec --scale 0.012 --variant split); run "powerlaw120k split" X=1
WL=(--workload c3_spec --scale 0.012 --variant split --half); run "powerlaw120k_f16 split" X=1
WL=(--workload c3_spec --scale 0.012 --half); run "powerlaw120k_f16 auto" X=1
WL=(--workload c3_spec --scale 0.05 ); run "powerlaw500k auto" X=1
WL=(--workload c3_spec --scale 0.05 --variant split); run "powerlaw500k split" X=1
tail -3 $OUT/small.err
echo done
