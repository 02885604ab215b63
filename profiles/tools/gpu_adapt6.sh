#!/bin/bash
TAG=${1:-ad6}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1500 python -m pytest tests/test_gpu_adaptive.py tests/test_golden.py tests/test_reference_golden.py -x -q -m gpu -k "adaptive" ) > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log | cut -c1-200
timeout 600 python profiles/tools/adaptive_time.py 2>$OUT/err | tee $OUT/adaptive.txt | grep "synthetic\|moon" | cut -c1-200
timeout 300 python profiles/tools/adapt_trace.py 2>>$OUT/err | tee $OUT/trace.txt
tail -3 $OUT/err
