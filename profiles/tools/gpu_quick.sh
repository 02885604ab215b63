#!/bin/bash
# quick check: GPU tests + one bench line
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests -x -q -m gpu ) > $OUT/pytest_gpu.log 2>&1
tail -12 $OUT/pytest_gpu.log
timeout 400 python bench.py --no-cpu > $OUT/bench.json 2> $OUT/bench.err; tail -3 $OUT/bench.err; cut -c1-1500 $OUT/bench.json
