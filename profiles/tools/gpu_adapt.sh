#!/bin/bash
# Development call for the adaptive rows: new tests without -x (all failures at once), then the rest of the GPU suite, then one bench line
TAG=${1:-adapt}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_adaptive.py -q -m gpu ) > $OUT/pytest_adapt.log 2>&1
tail -60 $OUT/pytest_adapt.log | cut -c1-400
( time timeout 900 python -m pytest tests -x -q -m gpu --ignore=tests/test_gpu_adaptive.py ) > $OUT/pytest_gpu.log 2>&1
tail -8 $OUT/pytest_gpu.log
timeout 400 python bench.py --no-cpu > $OUT/bench.json 2> $OUT/bench.err; tail -3 $OUT/bench.err; cut -c1-1200 $OUT/bench.json
