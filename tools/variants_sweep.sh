#!/bin/bash
# Per-config, per-variant timing (the MMA-vs-CUDA-core evidence north_star asks for). Writes JSON lines.
out=${1:-gpurun_out/variants.jsonl}
: > $out
for w in c4 c3 c5; do
  for v in auto cuda mma tma; do
    python bench.py --workload $w --variant $v --no-secondary --no-cpu --steps 30 --breakdown 2>/dev/null | tail -1 >> $out
  done
done
for w in c1 c2; do
  for v in auto cuda mma split; do
    python bench.py --workload $w --variant $v --no-secondary --no-cpu --steps 2000 --warmup 200 2>/dev/null | tail -1 >> $out
  done
  python bench.py --workload $w --cold --no-secondary --no-cpu --steps 200 --warmup 20 2>/dev/null | tail -1 >> $out
done
