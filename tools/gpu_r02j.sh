#!/bin/bash
OUT=gpurun_out/r02j
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
for lib in "" tools/libvariant_copies32.so tools/libvariant_part64k.so; do
  for s in "" "--slab 0/2"; do
    DASP_B200_LIB=${lib:+$PWD/$lib} timeout 300 python bench.py --workload c5_spec $s --steps 20 --warmup 5 $B --breakdown | sed "s|^{|{\"lib\": \"$lib\", \"slab\": \"$s\", |" >> $OUT/tune.jsonl 2>> $OUT/tune.err
  done
  DASP_B200_LIB=${lib:+$PWD/$lib} timeout 300 python bench.py --workload c3_spec --steps 20 --warmup 5 $B --breakdown | sed "s|^{|{\"lib\": \"$lib\", \"slab\": \"\", |" >> $OUT/tune.jsonl 2>> $OUT/tune.err
done
echo done
