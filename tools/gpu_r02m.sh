#!/bin/bash
# round 2, GPU call M (8 GPUs): the default bench at N=8 exactly as the driver launches it (+ reference arm)
OUT=gpurun_out/r02m
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
/usr/bin/time -v timeout 900 $TR --master-port 29561 bench.py --gpus 8 --steps 30 --warmup 5 > $OUT/bench_n8.json 2> $OUT/bench_n8.err
echo "bench n8 rc=$?"
timeout 400 $TR --master-port 29562 bench.py --gpus 8 --workload c5_spec --power-iter 100 --exchange hybrid --no-secondary --no-cpu --no-others > $OUT/power_hybrid_n8.json 2> $OUT/power_hybrid_n8.err
echo done
