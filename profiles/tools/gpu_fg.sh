#!/bin/bash
TAG=${1:-fg1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1500 python -m pytest tests -x -q -m gpu ) > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log | cut -c1-200
timeout 300 python profiles/tools/evaluators_time.py 2>$OUT/err | tee $OUT/evaluators.txt | cut -c1-200
tail -3 $OUT/err
