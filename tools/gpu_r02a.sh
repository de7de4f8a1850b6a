#!/bin/bash
# round 2, GPU call A: GPU tests, default bench, variant sweep on both generators of C3 / C5
OUT=gpurun_out/r02a
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total --format=csv > $OUT/gpu.txt 2>&1
free -g > $OUT/host_mem.txt; nproc >> $OUT/host_mem.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log
tail -5 $OUT/pytest.log
timeout 900 python bench.py --steps 30 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err
echo "bench rc=$?"
for w in c3_spec c3 c5_spec c5; do
  for v in auto cuda blocked; do
    timeout 300 python bench.py --workload $w --variant $v --steps 20 --warmup 5 --no-secondary --no-cpu --no-others --breakdown >> $OUT/sweep.jsonl 2>> $OUT/sweep.err
  done
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
echo done
