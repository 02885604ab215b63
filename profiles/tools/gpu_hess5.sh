#!/bin/bash
TAG=${1:-hess5}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests -x -q -m gpu -k "hess or shims or unregistered or trust or golden or anchor or adaptive" ) > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log | cut -c1-200
timeout 300 python profiles/tools/evaluators_time.py 2>$OUT/err | grep "hess_l\|grad" | cut -c1-260
echo "--- MPX_HESS_ROWS=0"
MPX_HESS_ROWS=0 timeout 300 python profiles/tools/evaluators_time.py 2>>$OUT/err | grep hess_l | cut -c1-260
timeout 300 python profiles/tools/hess_trace.py 2>>$OUT/err | tee $OUT/trace_rows.txt
MPX_HESS_ROWS=0 timeout 300 python profiles/tools/hess_trace.py 2>>$OUT/err | tee $OUT/trace_scattered.txt
tail -3 $OUT/err
