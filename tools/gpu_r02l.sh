#!/bin/bash
# round 2, GPU call L (2 GPUs): hybrid 1.5-D iterated workload: tests + timing at N=2
OUT=gpurun_out/r02l
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 -p no:cacheprovider -x > $OUT/pytest_multi.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_multi.log; tail -15 $OUT/pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29552 bench.py --gpus 2 --workload c5_spec --power-iter 100 --exchange hybrid --no-secondary --no-cpu --no-others > $OUT/power_hybrid_n2.json 2> $OUT/power_hybrid_n2.err
timeout 400 python bench.py --workload c5_spec --power-iter 100 --exchange hybrid --no-secondary --no-cpu --no-others > $OUT/power_hybrid_n1.json 2> $OUT/power_hybrid_n1.err
echo done
