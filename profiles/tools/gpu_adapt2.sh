#!/bin/bash
TAG=${1:-adapt2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_adaptive.py -q -m gpu -k "h_adaptive or mpopt_adaptive" ) > $OUT/pytest_adapt.log 2>&1
tail -40 $OUT/pytest_adapt.log | cut -c1-300
timeout 300 python profiles/tools/adaptive_time.py > $OUT/adaptive_time.txt 2> $OUT/adaptive_time.err; cat $OUT/adaptive_time.txt; tail -5 $OUT/adaptive_time.err
