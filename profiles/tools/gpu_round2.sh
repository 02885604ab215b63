#!/bin/bash
# GPU suite + evaluator timings + bench (development call)
TAG=${1:-round2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $OUT/pytest_gpu.log 2>&1
tail -8 $OUT/pytest_gpu.log | cut -c1-300
timeout 300 python profiles/tools/evaluators_time.py > $OUT/evaluators.txt 2> $OUT/evaluators.err; cat $OUT/evaluators.txt; tail -3 $OUT/evaluators.err
MPX_FGRAD_FUSED=0 timeout 300 python profiles/tools/evaluators_time.py 2>/dev/null | grep "f + grad" > $OUT/evaluators_fgrad_two_launches.txt; cat $OUT/evaluators_fgrad_two_launches.txt
timeout 300 python profiles/tools/adaptive_time.py > $OUT/adaptive_time.txt 2> $OUT/adaptive_time.err; grep "f + grad" $OUT/adaptive_time.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
