#!/bin/bash
OUT=gpurun_out/r02zm
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
run() { tag=$1; shift; echo "# $tag" >> $OUT/small.jsonl; timeout 120 env "$@" python bench.py "${WL[@]}" --steps 1000 --warmup 100 $B --breakdown >> $OUT/small.jsonl 2>> $OUT/small.err; }
WL=(--workload c3_spec --scale 0.012); run "powerlaw120k auto" X=1; 
WL=(--workload c3_spec --scale 0.012 --variant cuda); run "powerlaw120k cuda(chunked long)" X=1
WL=(--workload c3_spec --scale 0.012 --variant blocked); run "powerlaw120k blocked" X=1
WL=(--workload c5_spec --scale 0.004); run "skewed200k auto" X=1
WL=(--workload c5_spec --scale 0.004 --variant blocked); run "skewed200k blocked" X=1
tail -3 $OUT/small.err
echo done
