#!/bin/bash
# round 2, GPU call G (2 GPUs): two-rank tests of the iterated workload, the default bench at N=2, SB v3 A/B, gather floor
OUT=gpurun_out/r02g
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
B="--no-secondary --no-cpu --no-others --no-iterated"
timeout 900 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_parity.py::test_two_handles_keep_their_devices" -m gpu -q --timeout 600 -p no:cacheprovider > $OUT/pytest_multi.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_multi.log; tail -4 $OUT/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err
echo "bench n2 rc=$?"
for v in auto nobands; do
  timeout 300 python bench.py --workload c5_spec --variant $v --steps 20 --warmup 5 $B --breakdown >> $OUT/sweep.jsonl 2>> $OUT/sweep.err
done
timeout 300 python bench.py --workload c3_spec --variant banded --steps 20 --warmup 5 $B --breakdown >> $OUT/sweep.jsonl 2>> $OUT/sweep.err
timeout 120 tools/gather_bench > $OUT/gather_bench.txt 2>&1
du -sh $OUT; echo done
