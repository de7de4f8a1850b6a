#!/bin/bash
# round 2, GPU call U: full GPU suite after the allocator / pack_reg changes, dasp_create timing, default bench + reference arm
OUT=gpurun_out/r02u
mkdir -p $OUT
timeout 300 python tools/preprocess_time.py 256 >> $OUT/preprocess_time.txt 2>&1
timeout 300 python tools/preprocess_time.py 256 DASP_TRACE_PREPROCESS=1 >> $OUT/preprocess_time.txt 2>&1
cat $OUT/preprocess_time.txt
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider ) > $OUT/pytest_all.log 2>&1
tail -6 $OUT/pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > $OUT/bench_reference.json 2> $OUT/bench_reference.err
( time timeout 1200 python bench.py --steps 20 --warmup 5 ) > $OUT/bench_default.json 2> $OUT/bench_default.err
echo "bench rc=$?"; tail -4 $OUT/bench_default.err
echo done
