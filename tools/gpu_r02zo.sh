#!/bin/bash
# round 2, GPU call ZO: final defaults: full GPU suite, smoke, default bench, small-matrix spot checks
OUT=gpurun_out/r02zo
mkdir -p $OUT
B="--no-secondary --no-cpu --no-others --no-iterated"
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x ) > $OUT/pytest_all.log 2>&1
tail -4 $OUT/pytest_all.log | head -2
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
( time timeout 1200 python bench.py --steps 20 --warmup 5 ) > $OUT/bench_default.json 2> $OUT/bench_default.err
echo "bench rc=$?"
run() { tag=$1; shift; echo "# $tag" >> $OUT/small.jsonl; timeout 120 env "$@" python bench.py "${WL[@]}" --steps 1000 --warmup 100 $B >> $OUT/small.jsonl 2>> $OUT/small.err; }
WL=(--workload c4 --grid 40); run "stencil40 auto" X=1
WL=(--workload c1); run "c1 auto" X=1
WL=(--workload c2); run "c2 auto" X=1
echo done
