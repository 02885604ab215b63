#!/bin/bash
TAG=${1:-tests}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 1200 python -m pytest tests -q -m gpu ) > $OUT/pytest_gpu.log 2>&1
tail -40 $OUT/pytest_gpu.log | cut -c1-300
