#!/bin/bash
TAG=${1:-hess2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests -x -q -m gpu -k "hess or shims or unregistered or trust" ) > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log | cut -c1-200
timeout 300 python profiles/tools/evaluators_time.py 2>/dev/null | grep "hess_l\|f + grad"
