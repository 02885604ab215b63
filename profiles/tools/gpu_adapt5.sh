#!/bin/bash
TAG=${1:-ad5}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1500 python -m pytest tests/test_gpu_adaptive.py tests/test_golden.py -x -q -m gpu ) > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log | cut -c1-200
timeout 600 python profiles/tools/adaptive_time.py 2>$OUT/err | tee $OUT/adaptive.txt | grep "synthetic\|moon" | grep -v hess | cut -c1-200
for c in 5; do echo "MPX_ADAPT_CTAS=$c"; MPX_ADAPT_CTAS=$c timeout 600 python profiles/tools/adaptive_time.py 2>>$OUT/err | grep "synthetic.*g + jac_g" | cut -c1-160; done
echo "MPX_QUEUE=0 (one CTA per segment)"; MPX_QUEUE=0 timeout 600 python profiles/tools/adaptive_time.py 2>>$OUT/err | grep "synthetic.*g + jac_g" | cut -c1-160
timeout 300 python profiles/tools/adapt_trace.py 2>>$OUT/err | tee $OUT/trace.txt
tail -3 $OUT/err
