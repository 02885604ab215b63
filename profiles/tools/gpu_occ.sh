#!/bin/bash
OUT=gpurun_out/$1
mkdir -p $OUT
for nb in 2 1; do echo "== MPX_ADAPT_NBUF=$nb (80 registers)"; MPX_ADAPT_NBUF=$nb timeout 300 python profiles/tools/adaptive_time.py 2>/dev/null | grep "g + jac_g" | cut -c1-120; done
( timeout 600 python -m pytest tests/test_gpu_adaptive.py -x -q -m gpu -k "matches_oracle" ) 2>&1 | tail -2
( MPX_ADAPT_NBUF=1 timeout 600 python -m pytest tests/test_gpu_adaptive.py -x -q -m gpu -k "matches_oracle" ) 2>&1 | tail -2
