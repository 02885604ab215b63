#!/bin/bash
TAG=${1:-hess}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mpx_hess_kernel -s 3 -c 1 -f -o $OUT/prof_hess python profiles/tools/evaluators_time.py > $OUT/ncu_hess.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mpx_fgrad_kernel -s 3 -c 1 -f -o $OUT/prof_fgrad python profiles/tools/evaluators_time.py > $OUT/ncu_fgrad.log 2>&1
ls -la $OUT
