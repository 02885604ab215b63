#!/bin/bash
TAG=${1:-resid}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests -x -q -m gpu -k "residual or h_adaptive or interpol or mp_api or golden" ) > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log | cut -c1-200
timeout 300 python profiles/tools/evaluators_time.py 2>/dev/null | tail -1
