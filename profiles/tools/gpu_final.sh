#!/bin/bash
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > $OUT/pytest_gpu.log 2>&1
tail -4 $OUT/pytest_gpu.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mpx_adapt_kernel -s 1 -c 1 -f -o $OUT/prof_adapt python profiles/tools/adaptive_probe.py 3 > $OUT/ncu_adapt.log 2>&1
ls -la $OUT
