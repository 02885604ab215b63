#!/bin/bash
# last call of the round: the driver's own sequence (GPU tests, smoke, bench, reference arm) + evaluator timings
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > $OUT/pytest_gpu.log 2>&1
tail -4 $OUT/pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 400 python bench.py > $OUT/bench.json 2> $OUT/bench.err; cut -c1-330 $OUT/bench.json; tail -2 $OUT/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-200 $OUT/bench_ref.json
timeout 300 python profiles/tools/evaluators_time.py > $OUT/evaluators.txt 2> $OUT/evaluators.err; cat $OUT/evaluators.txt; tail -2 $OUT/evaluators.err
