#!/bin/bash
OUT=gpurun_out/$1
mkdir -p $OUT
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log | cut -c1-160
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py --no-cpu --steps 100 | cut -c1-200
