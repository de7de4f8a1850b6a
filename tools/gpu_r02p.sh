#!/bin/bash
# round 2, GPU call P (8 GPUs): the driver's own scaling command at N=8 (default legs: C4 strong scaling, e2e, iterated C5)
OUT=gpurun_out/r02p
mkdir -p $OUT
N=${1:-8}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 ) > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "rc=$?"; tail -3 $OUT/bench_n$N.err; head -c 600 $OUT/bench_n$N.json
