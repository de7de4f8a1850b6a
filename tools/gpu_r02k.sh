#!/bin/bash
# round 2, GPU call K (8 GPUs): the default bench at N=8 (C4 strong scaling, e2e, iterated C5), exchange modes of the iterated workload
OUT=gpurun_out/r02k
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt; nvidia-smi topo -m > $OUT/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench_n8.json 2> $OUT/bench_n8.err
echo "bench n8 rc=$?"
for ex in perm bcast mcu; do
  timeout 400 $TR --master-port 29542 bench.py --gpus 8 --workload c5_spec --power-iter 100 --exchange $ex --no-secondary --no-cpu --no-others >> $OUT/power_n8.jsonl 2>> $OUT/power_n8.err
done
timeout 400 $TR --master-port 29543 bench.py --gpus 8 --workload c5 --power-iter 100 --exchange perm --no-secondary --no-cpu --no-others >> $OUT/power_n8.jsonl 2>> $OUT/power_n8.err
echo done
